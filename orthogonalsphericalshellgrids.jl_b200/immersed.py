"""Fold-consistent bottom height for immersed-boundary grids over a tripolar grid.

The reference has no immersed-boundary code of its own: it only widens its grid aliases so that an
`ImmersedBoundaryGrid` wrapping a `TripolarGrid` takes the same boundary conditions
(`TRG`, src/tripolar_grid.jl:371; `DTRG`, src/distributed_tripolar_grid.jl:15).  What makes the
wrapped grid consistent is the zipper fill of the `(Center, Center, Nothing)` bottom-height field:
Oceananigans' `ImmersedBoundaryGrid(grid, GridFittedBottom(h))` sets the field from `h(λ, φ)` and
calls `fill_halo_regions!`, and the Center-Center zipper overwrites the right half of row Ny with
the mirror of the left half (src/zipper_boundary_condition.jl:122-138), so both images of every
cell on the fold line see one depth (test/test_zipper_boundary_conditions.jl:48-54).  This module
mirrors exactly that much: the two constructors, the filled bottom-height field and the resulting
cell mask.  The fill runs in libossg_b200.so like any other field's.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .fields import Field, fill_halo_regions, set_
from .grids import Center, OrthogonalSphericalShellGrid


@dataclass
class GridFittedBottom:
    """bottom_height: a number, an array of shape (Ny, Nx), a function (λ, φ) -> height evaluated
    at the Center-Center nodes, or (after construction of the grid) the filled Field."""
    bottom_height: object


class ImmersedBoundaryGrid:
    """ImmersedBoundaryGrid(grid, GridFittedBottom(h)) over a (serial or distributed) tripolar
    grid.  Everything the underlying grid has is reachable on the wrapper, so fields, boundary
    conditions and fills treat both alike (the reference's TRG / DTRG unions)."""

    def __init__(self, grid: OrthogonalSphericalShellGrid, ib: GridFittedBottom):
        if isinstance(grid, ImmersedBoundaryGrid):
            raise ValueError("ImmersedBoundaryGrid needs an underlying (non-immersed) grid")
        self.underlying_grid = grid
        h = Field((Center, Center, None), grid)
        value = ib.bottom_height.interior() if isinstance(ib.bottom_height, Field) else ib.bottom_height
        set_(h, value)
        fill_halo_regions(h)
        self.immersed_boundary = GridFittedBottom(h)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "underlying_grid"), name)

    def __repr__(self):
        return f"ImmersedBoundaryGrid on {self.underlying_grid!r} with GridFittedBottom"


def immersed_cell(grid: ImmersedBoundaryGrid) -> torch.Tensor:
    """Boolean mask over the padded (Nz, Ny+2Hy, Nx+2Hx) cells: True where the cell centre lies at
    or below the bottom (GridFittedBottom's centre criterion, z_center <= bottom_height)."""
    h = grid.immersed_boundary.bottom_height.data[0]
    zc = torch.as_tensor(grid.z.zc, dtype=h.dtype, device=h.device)
    return zc[:, None, None] <= h[None, :, :]
