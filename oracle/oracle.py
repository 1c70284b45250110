"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
leg may import this module.  It wraps oracle/libossg_oracle.so (own deterministic math)
and oracle/libossg_oracle_libm.so (glibc math, literal per-point evaluation), both built
from oracle/ossg_oracle.c by oracle/Makefile.

Arrays are returned as numpy arrays of shape (Ny+2Hy, Nx+2Hx) (C order) which is the
reference's column-major (Nx+2Hx, Ny+2Hy) parent array; use :func:`at` for Julia-style
1-based offset indexing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

COORD_NAMES = ["lam_cc", "lam_fc", "lam_cf", "lam_ff", "phi_cc", "phi_fc", "phi_cf", "phi_ff"]
METRIC_NAMES = ["dx_cc", "dx_fc", "dx_cf", "dx_ff", "dy_cc", "dy_cf", "dy_fc", "dy_ff",
                "az_cc", "az_fc", "az_cf", "az_ff"]
ALL_NAMES = COORD_NAMES + METRIC_NAMES
# (locx, locy) of each of the 20 arrays, 0 = Center, 1 = Face
LOC = {"cc": (0, 0), "fc": (1, 0), "cf": (0, 1), "ff": (1, 1)}
MATH_FN = {"sind": 0, "cosd": 1, "tand": 2, "sin": 3, "cos": 4, "atan": 5, "asin": 6,
           "asinh": 7, "sinh": 8, "cosh": 9}


class Params(C.Structure):
    _fields_ = [("Nx", C.c_int), ("Ny", C.c_int), ("Hx", C.c_int), ("Hy", C.c_int),
                ("southernmost_latitude", C.c_double), ("north_poles_latitude", C.c_double),
                ("first_pole_longitude", C.c_double), ("radius", C.c_double)]


R_EARTH = 6371.0e3


def make_params(size, halo=(4, 4, 4), southernmost_latitude=-80, north_poles_latitude=55,
                first_pole_longitude=70, radius=R_EARTH) -> Params:
    return Params(int(size[0]), int(size[1]), int(halo[0]), int(halo[1]),
                  float(southernmost_latitude), float(north_poles_latitude),
                  float(first_pole_longitude), float(radius))


def build_libraries(force: bool = False) -> None:
    """(Re)build the oracle shared objects with gcc (seconds)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n))
                            for n in ("libossg_oracle.so", "libossg_oracle_libm.so"))
    if not need:
        srcs = [os.path.join(_HERE, n) for n in ("ossg_oracle.c", "ossg_oracle.h", "ossg_oracle_math.h", "Makefile")]
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(os.path.getmtime(os.path.join(_HERE, n)) < newest
                   for n in ("libossg_oracle.so", "libossg_oracle_libm.so"))
    if need:
        subprocess.run(["make", "-C", _HERE, "CC=gcc", "-B"], check=True, capture_output=True)


class Oracle:
    def __init__(self, libm: bool = False):
        build_libraries()
        name = "libossg_oracle_libm.so" if libm else "libossg_oracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        L = self.lib
        dpp = C.POINTER(C.c_void_p)
        L.ossg_oracle_check_params.argtypes = [C.POINTER(Params)]
        L.ossg_oracle_coordinate_vectors.argtypes = [C.POINTER(Params)] + [C.c_void_p] * 4
        L.ossg_oracle_coordinate_vectors.restype = None
        L.ossg_oracle_coords.argtypes = [C.POINTER(Params), dpp]
        L.ossg_oracle_metrics.argtypes = [C.POINTER(Params), dpp, dpp]
        L.ossg_oracle_build_f64.argtypes = [C.POINTER(Params), dpp]
        L.ossg_oracle_build_f32.argtypes = [C.POINTER(Params), dpp]
        for suffix in ("f64", "f32"):
            z = getattr(L, f"ossg_oracle_zipper_{suffix}")
            z.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_double, C.c_int, C.c_int]
            z.restype = None
            pz = getattr(L, f"ossg_oracle_periodic_x_{suffix}")
            pz.argtypes = [C.c_void_p] + [C.c_int] * 6
            pz.restype = None
        L.ossg_oracle_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ossg_oracle_set_threads.argtypes = [C.c_int]
        L.ossg_oracle_set_threads.restype = None
        assert bool(L.ossg_oracle_uses_libm()) == libm

    # -- helpers
    @staticmethod
    def _ptrs(arrays):
        arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        return C.cast(arr, C.POINTER(C.c_void_p))

    @staticmethod
    def padded_shape(p: Params):
        return (p.Ny + 2 * p.Hy, p.Nx + 2 * p.Hx)

    def set_threads(self, n: int) -> None:
        self.lib.ossg_oracle_set_threads(int(n))

    def set_lambda_f32(self, on: bool) -> None:
        """FT = Float32 reading of generate_coordinate: 1-D longitudes rounded to Float32 before the map."""
        self.lib.ossg_oracle_set_lambda_f32.restype = None
        self.lib.ossg_oracle_set_lambda_f32(int(bool(on)))

    def max_threads(self) -> int:
        return int(self.lib.ossg_oracle_max_threads())

    # -- the path
    def coordinate_vectors(self, p: Params):
        lamF, lamC = np.empty(p.Nx), np.empty(p.Nx)
        phiF, phiC = np.empty(p.Ny), np.empty(p.Ny)
        self.lib.ossg_oracle_coordinate_vectors(C.byref(p), lamF.ctypes.data, lamC.ctypes.data,
                                                phiF.ctypes.data, phiC.ctypes.data)
        return lamF, lamC, phiF, phiC

    def coords(self, p: Params) -> dict:
        out = [np.empty(self.padded_shape(p)) for _ in range(8)]
        rc = self.lib.ossg_oracle_coords(C.byref(p), self._ptrs(out))
        if rc:
            raise ValueError(f"oracle: invalid parameters ({rc})")
        return dict(zip(COORD_NAMES, out))

    def metrics(self, p: Params, coords: dict) -> dict:
        cin = [np.ascontiguousarray(coords[n], dtype=np.float64) for n in COORD_NAMES]
        out = [np.empty(self.padded_shape(p)) for _ in range(12)]
        rc = self.lib.ossg_oracle_metrics(C.byref(p), self._ptrs(cin), self._ptrs(out))
        if rc:
            raise ValueError(f"oracle: invalid parameters ({rc})")
        return dict(zip(METRIC_NAMES, out))

    def build(self, p: Params, dtype=np.float64) -> dict:
        out = [np.empty(self.padded_shape(p), dtype=dtype) for _ in range(20)]
        fn = self.lib.ossg_oracle_build_f64 if dtype == np.float64 else self.lib.ossg_oracle_build_f32
        rc = fn(C.byref(p), self._ptrs(out))
        if rc:
            raise ValueError(f"oracle: invalid parameters ({rc})")
        return dict(zip(ALL_NAMES, out))

    def zipper(self, c: np.ndarray, Nx, Ny, Nz, Hx, Hy, Hz, locx, locy, sign, k0=None, nk=None):
        """In-place north fold of the parent array c[(Nz+2Hz), (Ny+2Hy), (Nx+2Hx)]."""
        assert c.flags.c_contiguous and c.size == (Nx + 2 * Hx) * (Ny + 2 * Hy) * (Nz + 2 * Hz)
        k0 = Hz if k0 is None else k0
        nk = Nz if nk is None else nk
        fn = {np.dtype(np.float64): self.lib.ossg_oracle_zipper_f64,
              np.dtype(np.float32): self.lib.ossg_oracle_zipper_f32}[c.dtype]
        fn(c.ctypes.data, Nx, Ny, Nz, Hx, Hy, Hz, locx, locy, float(sign), k0, nk)
        return c

    def periodic_x(self, c: np.ndarray, Nx, Ny, Nz, Hx, Hy, Hz):
        assert c.flags.c_contiguous
        fn = {np.dtype(np.float64): self.lib.ossg_oracle_periodic_x_f64,
              np.dtype(np.float32): self.lib.ossg_oracle_periodic_x_f32}[c.dtype]
        fn(c.ctypes.data, Nx, Ny, Nz, Hx, Hy, Hz)
        return c

    def fill_halo_regions(self, c, Nx, Ny, Nz, Hx, Hy, Hz, locx, locy, sign):
        """zipper north (interior k-levels) then periodic x over the whole parent extent."""
        self.zipper(c, Nx, Ny, Nz, Hx, Hy, Hz, locx, locy, sign)
        return self.periodic_x(c, Nx, Ny, Nz, Hx, Hy, Hz)

    def cartesian_nodes(self, lam, phi, radius=1.0):
        """(external) Oceananigans.Grids.lat_lon_to_cartesian over arrays of nodes."""
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        out = [np.empty_like(lam) for _ in range(3)]
        self.lib.ossg_oracle_cartesian_nodes.restype = None
        self.lib.ossg_oracle_cartesian_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ossg_oracle_cartesian_nodes(lam.ctypes.data, phi.ctypes.data, lam.size, float(radius), *[o.ctypes.data for o in out])
        return out

    def math(self, name: str, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        rc = self.lib.ossg_oracle_math_eval(MATH_FN[name], x.ctypes.data, y.ctypes.data, x.size)
        assert rc == 0
        return y


def at(a: np.ndarray, p: Params, i, j):
    """Julia-style a[i, j] on a padded 2-D array (1-based, halos at i<1, j<1, ...)."""
    return a[j + p.Hy - 1, i + p.Hx - 1]


def interior(a: np.ndarray, p: Params) -> np.ndarray:
    """Interior view, shape (Ny, Nx)."""
    return a[p.Hy:p.Hy + p.Ny, p.Hx:p.Hx + p.Nx]


@dataclass
class ReadmeSummary:
    centre: tuple
    lon_extent: float
    lat_extent: float
    dlam_min: float
    dlam_max: float
    dphi_min: float
    dphi_max: float


def readme_summary(g: dict, p: Params) -> ReadmeSummary:
    """The six numbers Oceananigans prints for the grid (README.md:54-59), in degrees."""
    R = p.radius
    deg = 180.0 / np.pi
    ic, jc = p.Nx // 2 + 1, p.Ny // 2 + 1
    dxcf = interior(g["dx_cf"], p)
    dyfc = interior(g["dy_fc"], p)
    dxff = interior(g["dx_ff"], p)
    dyff = interior(g["dy_ff"], p)
    return ReadmeSummary(
        centre=(float(at(g["lam_ff"], p, ic, jc)), float(at(g["phi_ff"], p, ic, jc))),
        lon_extent=float(dxcf.sum(axis=1).max() / R * deg),
        lat_extent=float(dyfc.sum(axis=0).max() / R * deg),
        dlam_min=float(dxff.min() / R * deg), dlam_max=float(dxff.max() / R * deg),
        dphi_min=float(dyff.min() / R * deg), dphi_max=float(dyff.max() / R * deg))
