"""Host-side mirror of the reference's field / boundary-condition interface for the fold.

Mirrors:
  ZipperBoundaryCondition(sign = 1), Zipper, ZBC          src/zipper_boundary_condition.jl:8,52,54
  validate_boundary_condition_location                    src/zipper_boundary_condition.jl:58-62
  regularize_field_boundary_conditions(bcs, grid, name)   src/tripolar_grid_extensions.jl:25-44
  sign(LX, LY)                                            src/tripolar_grid_extensions.jl:49-53
  Field((LX, LY, LZ), grid, ...)                          src/tripolar_grid_extensions.jl:57-80
  fill_halo_regions!(field) -> _fill_north_halo!          src/zipper_boundary_condition.jl:146-155
The fill itself runs in libossg_b200.so (ossg_fill_halo_regions_batch / ossg_zipper_fill_batch).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import FieldDesc
from .grids import Center, Face, OrthogonalSphericalShellGrid, _loc_code, _require_cuda


class Zipper:
    """Boundary-condition classification (src/zipper_boundary_condition.jl:8)."""

    def __repr__(self):
        return "Zipper"


class Periodic:
    def __repr__(self):
        return "Periodic"


@dataclass(frozen=True)
class BoundaryCondition:
    classification: object
    condition: object = None


def ZipperBoundaryCondition(sign=1) -> BoundaryCondition:
    """src/zipper_boundary_condition.jl:52"""
    return BoundaryCondition(Zipper(), sign)


def PeriodicBoundaryCondition() -> BoundaryCondition:
    return BoundaryCondition(Periodic(), None)


def is_zipper(bc) -> bool:
    return isinstance(bc, BoundaryCondition) and isinstance(bc.classification, Zipper)


def bc_str(bc) -> str:
    """src/zipper_boundary_condition.jl:56"""
    return "Zipper" if is_zipper(bc) else repr(bc)


def validate_boundary_condition_location(bc, loc, side):
    """src/zipper_boundary_condition.jl:58-62: a Zipper is only valid on the north side."""
    if isinstance(bc, Zipper) or is_zipper(bc):
        if side != "north":
            raise ValueError(f"Cannot specify {side} boundary condition Zipper on a field at {loc!r} (north only)!")
    return None


@dataclass
class FieldBoundaryConditions:
    west: object = None
    east: object = None
    south: object = None
    north: object = None
    bottom: object = None
    top: object = None
    immersed: object = None


def sign(LX, LY) -> int:
    """src/tripolar_grid_extensions.jl:49-53: fields on edges are signed vectors."""
    fx, fy = _loc_code(LX) == _lib.OSSG_FACE, _loc_code(LY) == _lib.OSSG_FACE
    return -1 if fx != fy else 1


_ASSUMED_LOCATION = {"u": (Face, Center, Center), "v": (Center, Face, Center), "w": (Center, Center, Face)}


def assumed_field_location(name: str):
    return _ASSUMED_LOCATION.get(name, (Center, Center, Center))


def regularize_field_boundary_conditions(bcs: FieldBoundaryConditions, grid, field_name: str,
                                         prognostic_names=None) -> FieldBoundaryConditions:
    """src/tripolar_grid_extensions.jl:25-44 (and the Distributed variant,
    src/distributed_tripolar_grid.jl:129-155): north is a Zipper with sign -1 for :u, :v."""
    from .distributed import Distributed
    s = -1 if field_name in ("u", "v") else 1
    north = ZipperBoundaryCondition(s)
    arch = grid.architecture
    if isinstance(arch, Distributed) and not arch.is_top_row():
        north = bcs.south       # sic: the reference regularises bcs.south for the north side (:146)
    return FieldBoundaryConditions(bcs.west, bcs.east, bcs.south, north, bcs.bottom, bcs.top, bcs.immersed)


def default_boundary_conditions() -> FieldBoundaryConditions:
    return FieldBoundaryConditions(west=PeriodicBoundaryCondition(), east=PeriodicBoundaryCondition())


class Field:
    """A 3-D field on a tripolar grid: parent tensor of shape (Nz+2Hz, Ny+2Hy, Nx+2Hx) on the
    grid's device.  If the supplied north boundary condition is not a Zipper it is replaced by
    ZipperBoundaryCondition(sign(LX, LY)) (src/tripolar_grid_extensions.jl:57-80).
    LZ = None is Julia's `Nothing` location: a horizontal field with one level and no z halo
    (parent shape (1, Ny+2Hy, Nx+2Hx)), e.g. the bottom height of an immersed boundary."""

    def __init__(self, loc, grid: OrthogonalSphericalShellGrid, data: Optional[torch.Tensor] = None,
                 boundary_conditions: Optional[FieldBoundaryConditions] = "default", dtype=None):
        from .distributed import Distributed
        LX, LY, LZ = loc
        self.loc = (LX, LY, LZ)
        self.grid = grid
        dtype = dtype or grid.dtype
        self.Nz, self.Hz = (1, 0) if LZ is None else (grid.Nz, grid.Hz)
        shape = (self.Nz + 2 * self.Hz, grid.Ny + 2 * grid.Hy, grid.Nx + 2 * grid.Hx)
        if data is None:
            device = grid.arrays["lam_cc"].parent.device
            data = torch.zeros(shape, dtype=dtype, device=device)
        elif tuple(data.shape) != shape:      # validate_field_data
            raise ValueError(f"data has shape {tuple(data.shape)}, the grid needs {shape}")
        self.data = data
        if boundary_conditions == "default":
            boundary_conditions = default_boundary_conditions()
        if boundary_conditions is not None:
            old = boundary_conditions
            for side in ("west", "east", "south", "bottom", "top"):
                validate_boundary_condition_location(getattr(old, side), LX, side)
            arch = grid.architecture
            top = not isinstance(arch, Distributed) or arch.is_top_row()
            if top:
                north = old.north if is_zipper(old.north) else ZipperBoundaryCondition(sign(LX, LY))
            else:
                north = old.north          # halo-communication BC injected by the distributed layer
            boundary_conditions = FieldBoundaryConditions(old.west, old.east, old.south, north,
                                                          old.bottom, old.top, old.immersed)
        self.boundary_conditions = boundary_conditions

    # -- reference-style helpers
    def interior(self) -> torch.Tensor:
        g = self.grid
        return self.data[self.Hz:self.Hz + self.Nz, g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx]

    def __getitem__(self, ijk):
        """Julia indexing c[i, j, k] (1-based interior, halos outside)."""
        g = self.grid
        i, j, k = ijk

        def ix(v, H):
            if isinstance(v, slice):
                lo = None if v.start is None else v.start + H - 1
                hi = None if v.stop is None else v.stop + H
                return slice(lo, hi)
            return v + H - 1
        return self.data[ix(k, self.Hz), ix(j, g.Hy), ix(i, g.Hx)]

    def descriptor(self) -> FieldDesc:
        g = self.grid
        bc = self.boundary_conditions.north if self.boundary_conditions is not None else None
        s = float(bc.condition) if is_zipper(bc) else 1.0
        et = _lib.OSSG_F64 if self.data.dtype == torch.float64 else _lib.OSSG_F32
        return FieldDesc(self.data.data_ptr(), et, g.Nx, g.Ny, self.Nz, g.Hx, g.Hy, self.Hz,
                         _loc_code(self.loc[0]), _loc_code(self.loc[1]), 0, 0, s)


def CenterField(grid, **kw):
    return Field((Center, Center, Center), grid, **kw)


def XFaceField(grid, **kw):
    return Field((Face, Center, Center), grid, **kw)


def YFaceField(grid, **kw):
    return Field((Center, Face, Center), grid, **kw)


def ZFaceField(grid, **kw):
    return Field((Center, Center, Face), grid, **kw)


_NODE = {(0, 0): "cc", (1, 0): "fc", (0, 1): "cf", (1, 1): "ff"}


def set_(f: Field, value) -> Field:
    """set!(field, value): a number, an array of the interior shape (Nz, Ny, Nx), or a function
    evaluated at the field's nodes: (λ, φ, z) for a 3-D field, (λ, φ) for a horizontal one."""
    g = f.grid
    if callable(value):
        key = _NODE[(_loc_code(f.loc[0]), _loc_code(f.loc[1]))]
        lam = g.arrays["lam_" + key].interior(g.Nx, g.Ny).to(f.data.dtype)
        phi = g.arrays["phi_" + key].interior(g.Nx, g.Ny).to(f.data.dtype)
        if f.loc[2] is None:
            v = value(lam[None, :, :], phi[None, :, :])
        else:
            z = g.z.zf if _loc_code(f.loc[2]) == _lib.OSSG_FACE else g.z.zc
            zt = torch.as_tensor(np.asarray(z)[:f.Nz], dtype=f.data.dtype, device=f.data.device)
            v = value(lam[None, :, :], phi[None, :, :], zt[:, None, None])
        v = torch.as_tensor(v, dtype=f.data.dtype, device=f.data.device)
        f.interior().copy_(v.expand(f.Nz, g.Ny, g.Nx))
    elif isinstance(value, (int, float)):
        f.interior().fill_(value)
    else:
        f.interior().copy_(torch.as_tensor(value, dtype=f.data.dtype, device=f.data.device))
    return f


def _batch(fields: Sequence[Field]):
    arr = (FieldDesc * len(fields))(*[f.descriptor() for f in fields])
    return arr


def fill_halo_regions(fields, stream: Optional[torch.cuda.Stream] = None, only_north: bool = False):
    """fill_halo_regions!(field) / tupled_fill_halo_regions! for fields on a (serial) tripolar grid:
    zipper north fill, then periodic west/east over the whole parent extent; one batched launch
    sequence for any number of fields.  Distributed grids go through distributed.fill_halo_regions."""
    from .distributed import Distributed, distributed_fill_halo_regions
    if isinstance(fields, Field):
        fields = [fields]
    fields = list(fields)
    if not fields:
        return
    if isinstance(fields[0].grid.architecture, Distributed):
        return distributed_fill_halo_regions(fields, stream=stream)
    _require_cuda()
    zips = [f for f in fields if f.boundary_conditions is not None and is_zipper(f.boundary_conditions.north)]
    if not zips:
        return
    dev = zips[0].data.device
    with torch.cuda.device(dev):
        s = torch.cuda.current_stream(dev) if stream is None else stream
        fn = _lib.lib().ossg_zipper_fill_batch if only_north else _lib.lib().ossg_fill_halo_regions_batch
        rc = fn(_batch(zips), len(zips), s.cuda_stream)
    _lib.check(rc)
