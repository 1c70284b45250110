"""ossg-b200: B200-native TripolarGrid construction and zipper halo fill
(drop-in for the hot path of CliMA/OrthogonalSphericalShellGrids.jl v0.2.1).

The reference exports two names, TripolarGrid and ZipperBoundaryCondition
(src/OrthogonalSphericalShellGrids.jl:4); the rest of this namespace mirrors the Oceananigans
generics the reference extends for the path."""
from ._lib import OSSGError, build as build_library, lib as load_library
from .grids import (cartesian_nodes, CPU, GPU, Center, Face, Float32, Float64, OffsetArray, OrthogonalSphericalShellGrid,
                    R_Earth, Tripolar, TripolarGrid, with_halo, x_domain, y_domain)
from .fields import (BoundaryCondition, CenterField, Field, FieldBoundaryConditions, PeriodicBoundaryCondition,
                     XFaceField, YFaceField, ZFaceField, Zipper, ZipperBoundaryCondition, bc_str,
                     fill_halo_regions, regularize_field_boundary_conditions, set_, sign,
                     validate_boundary_condition_location)
from .immersed import GridFittedBottom, ImmersedBoundaryGrid, immersed_cell
from .distributed import (Distributed, DistributedTripolarGrid, Partition, reconstruct_global_grid)

__all__ = ["TripolarGrid", "ZipperBoundaryCondition"]
