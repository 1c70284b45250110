"""Host-side mirror of the reference's grid interface for the hot path.

Mirrors (names, argument meaning, error behaviour):
  TripolarGrid(arch, FT; size, southernmost_latitude, halo, radius, z, north_poles_latitude,
               first_pole_longitude)                         src/tripolar_grid.jl:59-333
  Tripolar (conformal_mapping tag)                            src/tripolar_grid.jl:6-15
  with_halo(new_halo, grid)                                   src/with_halo.jl:5-23
Julia is not available in this environment, so the host side that would be a `ccall` shim is
written in Python over the same C symbols (INTEGRATION.md shows the Julia binding).  All
array work happens in libossg_b200.so on the GPU; torch is used only to own device memory.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field as _dc_field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ARRAY_NAMES, TripolarParams

R_Earth = 6371.0e3      # Oceananigans.Grids.R_Earth (external), src/tripolar_grid.jl:63

Float64 = torch.float64
Float32 = torch.float32


class Center:
    """Oceananigans.Center"""
    code = _lib.OSSG_CENTER

    def __repr__(self):
        return "Center"


class Face:
    """Oceananigans.Face"""
    code = _lib.OSSG_FACE

    def __repr__(self):
        return "Face"


def _loc_code(L) -> int:
    if L is Center or isinstance(L, Center):
        return _lib.OSSG_CENTER
    if L is Face or isinstance(L, Face):
        return _lib.OSSG_FACE
    if L is None:            # Nothing location (reduced dimension): treated as Center by the fold
        return _lib.OSSG_CENTER
    raise TypeError(f"not a location: {L!r}")


# ---- architectures -------------------------------------------------------------------------
class CPU:
    """Arrays are delivered in host memory.  The computation still runs on the GPU (the
    north-star has no CPU path); the 20 arrays are copied device -> host at the end."""

    def __repr__(self):
        return "CPU()"


class GPU:
    """Arrays stay on the current CUDA device (Oceananigans.GPU)."""

    def __init__(self, device: Optional[int] = None):
        self.device = device

    def torch_device(self) -> torch.device:
        idx = torch.cuda.current_device() if self.device is None else self.device
        return torch.device("cuda", idx)

    def __repr__(self):
        return "GPU()"


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this framework has no CPU fallback (TripolarGrid "
                           "coordinates/metrics and the zipper fill run only on sm_100a)")
    rc = _lib.lib().ossg_device_ok()
    if rc != 0:
        raise RuntimeError("unsupported device: " + _lib.last_error())


# ---- Julia-style offset view -----------------------------------------------------------------
class OffsetArray:
    """Parent tensor plus Julia OffsetArray indexing: ``a[i, j]`` with 1-based interior indices,
    halos at i < 1 etc.  The parent has shape (Ny+2Hy, Nx+2Hx) (C order) == the reference's
    column-major (Nx+2Hx, Ny+2Hy)."""

    def __init__(self, parent: torch.Tensor, Hx: int, Hy: int):
        self.parent = parent
        self.offsets = (-Hx, -Hy)
        self.Hx, self.Hy = Hx, Hy

    def _ix(self, i, H):
        if isinstance(i, slice):      # Julia a:b (inclusive) is written slice(a, b)
            lo = None if i.start is None else i.start + H - 1
            hi = None if i.stop is None else i.stop + H
            return slice(lo, hi)
        return i + H - 1

    def __getitem__(self, ij):
        i, j = ij
        return self.parent[self._ix(j, self.Hy), self._ix(i, self.Hx)]

    def interior(self, Nx: int, Ny: int) -> torch.Tensor:
        return self.parent[self.Hy:self.Hy + Ny, self.Hx:self.Hx + Nx]

    def __repr__(self):
        return f"OffsetArray({tuple(self.parent.shape)[::-1]}, offsets={self.offsets}, {self.parent.dtype}, {self.parent.device})"


@dataclass(frozen=True)
class Tripolar:
    """src/tripolar_grid.jl:6-10"""
    north_poles_latitude: float
    first_pole_longitude: float
    southernmost_latitude: float


@dataclass
class VerticalCoordinate:
    """1-D z coordinate (Oceananigans generate_coordinate(..., :z, 3, CPU()); host-side, not on
    the data-parallel path).  Faces zf[0..Nz], centres zc[0..Nz-1]."""
    zf: np.ndarray
    zc: np.ndarray
    Lz: float


def _generate_z(z, Nz: int) -> VerticalCoordinate:
    if callable(z):
        zf = np.array([z(k) for k in range(1, Nz + 2)], dtype=np.float64)
    else:
        zf = np.asarray(z, dtype=np.float64).ravel()
        if zf.size == 2:                      # z = (bottom, top): regularly spaced
            zf = np.linspace(zf[0], zf[1], Nz + 1)
        elif zf.size != Nz + 1:
            raise ValueError(f"z must be (bottom, top) or {Nz + 1} faces, got {zf.size} values")
    zc = 0.5 * (zf[1:] + zf[:-1])
    return VerticalCoordinate(zf, zc, float(zf[-1] - zf[0]))


@dataclass
class OrthogonalSphericalShellGrid:
    """Field-for-field counterpart of Oceananigans' OrthogonalSphericalShellGrid as filled by
    src/tripolar_grid.jl:304-330.  The 20 horizontal arrays are OffsetArrays over torch tensors;
    ASCII names lam_cc ... az_ff, with the reference's λᶜᶜᵃ-style names available as aliases."""
    architecture: object
    Nx: int
    Ny: int
    Nz: int
    Hx: int
    Hy: int
    Hz: int
    Lz: float
    z: VerticalCoordinate
    radius: float
    conformal_mapping: Tripolar
    dtype: torch.dtype
    topology: tuple = ("Periodic", "RightConnected", "Bounded")
    arrays: dict = _dc_field(default_factory=dict)
    # tile of a distributed grid (global index of the first interior cell, global size)
    i0: int = 1
    j0: int = 1
    global_size: Optional[tuple] = None
    # header + 1-D tables of the construction workspace (device); with_halo hands them to the next build
    tables: Optional[torch.Tensor] = None

    _ALIASES = {"λ": "lam", "φ": "phi", "Δx": "dx", "Δy": "dy", "Az": "az"}

    def __getattr__(self, name):
        arrays = object.__getattribute__(self, "arrays")
        if name in arrays:
            return arrays[name]
        # λᶜᶜᵃ / Δxᶠᶜᵃ ... (also after Python's NFKC identifier normalisation: λcca, Δxfca)
        import unicodedata
        n = unicodedata.normalize("NFKC", name)
        for pre, asc in self._ALIASES.items():
            if n.startswith(pre) and len(n) == len(pre) + 3 and n.endswith("a"):
                key = f"{asc}_{n[len(pre):len(pre) + 2]}"
                if key in arrays:
                    return arrays[key]
        raise AttributeError(name)

    @property
    def size(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def halo(self):
        return (self.Hx, self.Hy, self.Hz)

    def eltype(self):
        return self.dtype

    def parents(self) -> list:
        return [self.arrays[n].parent for n in ARRAY_NAMES]

    def __repr__(self):
        return (f"{self.Nx}×{self.Ny}×{self.Nz} OrthogonalSphericalShellGrid{{{self.dtype}, "
                f"{', '.join(self.topology)}}} on {self.architecture} with {self.Hx}×{self.Hy}×{self.Hz} halo "
                f"and with precomputed metrics, conformal_mapping = {self.conformal_mapping}")


# x_domain / y_domain, src/tripolar_grid_extensions.jl:20-21
def x_domain(grid):
    return (0, 360)


def y_domain(grid):
    return (float(grid.arrays["phi_ff"].parent.min()), 90)


# ---- workspace (device scratch: header + 1-D tables, plus the FP64 coordinates of Float32 grids) --
# Every build owns its workspace (two builds on different streams never share scratch).  The small
# table prefix stays alive with the grid (`grid.tables`): with_halo / reconstruct_global_grid hand
# it to the next build, whose table kernel then finds valid stamps and returns at once.
def _new_workspace(nbytes: int, device: torch.device, stream_ptr: int, tables: Optional[torch.Tensor]) -> torch.Tensor:
    L = _lib.lib()
    ws = torch.empty(max(int(nbytes), 128), dtype=torch.uint8, device=device)
    _lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), stream_ptr))
    if tables is not None and tables.device == ws.device:
        n = min(tables.numel(), ws.numel())
        ws[:n].copy_(tables[:n], non_blocking=True)        # stamps + tables of the earlier grid (same stream)
    return ws


def _elem_type(FT) -> int:
    if FT in (torch.float64, np.float64, float, "Float64"):
        return _lib.OSSG_F64
    if FT in (torch.float32, np.float32, "Float32"):
        return _lib.OSSG_F32
    raise TypeError(f"FT must be Float64 or Float32, got {FT!r}")


def make_params(size, halo, southernmost_latitude, north_poles_latitude, first_pole_longitude, radius,
                tile=None, flags: int = 0) -> TripolarParams:
    Nx, Ny = int(size[0]), int(size[1])
    i0, j0, nx, ny = tile if tile is not None else (1, 1, Nx, Ny)
    return TripolarParams(Nx, Ny, int(halo[0]), int(halo[1]), float(southernmost_latitude),
                          float(north_poles_latitude), float(first_pole_longitude), float(radius),
                          int(i0), int(j0), int(nx), int(ny), int(flags), 0)


def build_arrays_cached(params: TripolarParams, FT, device: torch.device, out: Optional[Sequence[torch.Tensor]] = None,
                        stream: Optional[torch.cuda.Stream] = None, tables: Optional[torch.Tensor] = None):
    """Run the construction kernels for the window described by `params`; returns (the 20 parent
    tensors of shape (ny+2Hy, nx+2Hx), the table prefix of the workspace to keep with the grid).
    Asynchronous on `stream` (default: current); allocations and copies are made on that stream too.
    `tables`: the table prefix of an earlier grid with the same size / latitudes (with_halo)."""
    L = _lib.lib()
    et = _elem_type(FT)
    tdt = torch.float64 if et == _lib.OSSG_F64 else torch.float32
    nx = params.nx or params.Nx
    ny = params.ny or params.Ny
    shape = (ny + 2 * params.Hy, nx + 2 * params.Hx)
    nbytes = L.ossg_tripolar_workspace_bytes(C.byref(params), et)
    if nbytes == 0:
        _lib.check(L.ossg_tripolar_build(C.byref(params), et, _lib.ptr_array([0] * 20), None, 0, None))
    with torch.cuda.device(device):
        s = torch.cuda.current_stream(device) if stream is None else stream
        with torch.cuda.stream(s):
            if out is None:
                out = [torch.empty(shape, dtype=tdt, device=device) for _ in range(20)]
            ws = _new_workspace(nbytes, device, s.cuda_stream, tables)
            rc = L.ossg_tripolar_build(C.byref(params), et, _lib.ptr_array([t.data_ptr() for t in out]),
                                       ws.data_ptr(), ws.numel(), s.cuda_stream)
            _lib.check(rc)
            tb = int(L.ossg_tripolar_table_bytes(C.byref(params)))
            # Float32 grids: the workspace also holds 8 FP64 coordinate arrays; keep only the table prefix
            keep = ws if ws.numel() <= 2 * tb else ws[:tb].clone()
    return list(out), keep


def build_arrays(params: TripolarParams, FT, device: torch.device, out: Optional[Sequence[torch.Tensor]] = None,
                 stream: Optional[torch.cuda.Stream] = None) -> list:
    """build_arrays_cached without the table prefix."""
    return build_arrays_cached(params, FT, device, out, stream)[0]


def TripolarGrid(arch=None, FT=Float64, *, size, southernmost_latitude=-80, halo=(4, 4, 4),
                 radius=R_Earth, z=(0, 1), north_poles_latitude=55, first_pole_longitude=70, _tables=None,
                 float32_longitudes: bool = False):
    """Return an OrthogonalSphericalShellGrid tripolar grid on the sphere
    (src/tripolar_grid.jl:59-333).  `arch` defaults to GPU(); a Distributed architecture
    dispatches to the distributed constructor (src/distributed_tripolar_grid.jl:24).

    float32_longitudes (not a reference keyword): the reference passes FT to generate_coordinate
    (:89), so for FT = Float32 its 1-D longitudes may be Float32 values.  True applies that reading
    (longitudes rounded once to Float32, OSSG_LAMBDA1D_F32); the default keeps them Float64 for every
    FT.  DESIGN.md section 4 has the reasoning and what it changes."""
    from .distributed import Distributed, DistributedTripolarGrid   # late import: optional path
    if arch is None:
        arch = GPU()
    if isinstance(arch, Distributed):
        return DistributedTripolarGrid(arch, FT, size=size, southernmost_latitude=southernmost_latitude,
                                       halo=halo, radius=radius, z=z,
                                       north_poles_latitude=north_poles_latitude,
                                       first_pole_longitude=first_pole_longitude, _tables=_tables,
                                       float32_longitudes=float32_longitudes)
    Nx, Ny, Nz = (int(s) for s in size)
    Hx, Hy, Hz = (int(h) for h in halo)
    if Nx % 2 == 1:          # src/tripolar_grid.jl:81-83
        raise ValueError("The number of cells in the longitude dimension should be even!")
    params = make_params(size, halo, southernmost_latitude, north_poles_latitude, first_pole_longitude, radius,
                         flags=_lib.OSSG_LAMBDA1D_F32 if float32_longitudes else 0)
    _require_cuda()
    device = arch.torch_device() if isinstance(arch, GPU) else torch.device("cuda", torch.cuda.current_device())
    parents, tables = build_arrays_cached(params, FT, device, tables=_tables)
    if isinstance(arch, CPU):
        parents = _to_host(parents)               # one pinned staging buffer, one synchronisation
    zc = _generate_z(z, Nz)
    tdt = parents[0].dtype
    grid = OrthogonalSphericalShellGrid(
        architecture=arch, Nx=Nx, Ny=Ny, Nz=Nz, Hx=Hx, Hy=Hy, Hz=Hz, Lz=zc.Lz, z=zc, radius=float(radius),
        conformal_mapping=Tripolar(north_poles_latitude, first_pole_longitude, southernmost_latitude),
        dtype=tdt, arrays={n: OffsetArray(t, Hx, Hy) for n, t in zip(ARRAY_NAMES, parents)},
        global_size=(Nx, Ny, Nz), tables=tables)
    return grid


def _to_host(parents: Sequence[torch.Tensor]) -> list:
    """Device -> host copy of the 20 arrays for TripolarGrid(CPU()): all copies are queued
    asynchronously into ONE pinned allocation and the stream is synchronised once (20 pageable
    `.cpu()` calls would each stage through a bounce buffer and synchronise)."""
    n = parents[0].numel()
    host = torch.empty((len(parents),) + tuple(parents[0].shape), dtype=parents[0].dtype, pin_memory=True)
    for k, t in enumerate(parents):
        host[k].copy_(t, non_blocking=True)
    torch.cuda.current_stream(parents[0].device).synchronize()
    return [host[k] for k in range(len(parents))]


def with_halo(new_halo, old_grid: OrthogonalSphericalShellGrid):
    """Rebuild the grid with a different halo (src/with_halo.jl:5-23, 25-44): re-runs the
    constructor, as the reference does, on the old grid's 1-D tables (their stamps are still valid,
    so the table kernel returns at once)."""
    from .distributed import Distributed
    cm = old_grid.conformal_mapping
    kw = dict(z=old_grid.z.zf, halo=new_halo, north_poles_latitude=cm.north_poles_latitude,
              first_pole_longitude=cm.first_pole_longitude, southernmost_latitude=cm.southernmost_latitude,
              _tables=getattr(old_grid, "tables", None))      # fast path: the 1-D tables do not depend on the halo
    if isinstance(old_grid.architecture, Distributed):
        # src/with_halo.jl:25-44: global size from the local sizes; `radius` is not forwarded there
        return TripolarGrid(old_grid.architecture, old_grid.dtype, size=old_grid.global_size, **kw)
    return TripolarGrid(old_grid.architecture, old_grid.dtype, size=old_grid.size, radius=old_grid.radius, **kw)


def cartesian_nodes(grid: OrthogonalSphericalShellGrid, loc=(Face, Face), radius=None, dtype=None):
    """Cartesian coordinates (x, y, z) of the grid's nodes at horizontal location `loc`, over the
    whole parent arrays -- Oceananigans' get_cartesian_nodes_and_vertices / lat_lon_to_cartesian
    as used by test/test_tripolar_grid.jl:68 and examples/visualize_tripolar_grid.jl:39-42, computed
    on the GPU (ossg_cartesian_nodes).  Needs the FP64 coordinates, i.e. a Float64 grid."""
    import ctypes as C
    suffix = ("f" if loc[0] is Face else "c") + ("f" if loc[1] is Face else "c")
    lam, phi = grid.arrays["lam_" + suffix].parent, grid.arrays["phi_" + suffix].parent
    if lam.dtype != torch.float64:
        raise ValueError("cartesian_nodes needs a Float64 grid (the coordinates are read as FP64)")
    dtype = dtype or torch.float64
    out = [torch.empty_like(lam, dtype=dtype) for _ in range(3)]
    L = _lib.lib()
    stream = torch.cuda.current_stream(lam.device).cuda_stream
    _lib.check(L.ossg_cartesian_nodes(lam.data_ptr(), phi.data_ptr(), lam.numel(), float(grid.radius if radius is None else radius),
                                      _lib.OSSG_F64 if dtype == torch.float64 else _lib.OSSG_F32,
                                      out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stream))
    return tuple(OffsetArray(t, grid.Hx, grid.Hy) for t in out)
