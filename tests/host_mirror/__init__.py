"""ctypes wrapper of the TEST-ONLY host build of the product's per-point headers
(tests/host_mirror/host_mirror.cpp).  Lets the CPU test-suite compare the product's
formulas and index maps with the oracle bit-for-bit without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_mirror.cpp")
_SO = os.path.join(_HERE, "libhost_mirror.so")
_CSRC = os.path.join(_HERE, "..", "..", "orthogonalsphericalshellgrids.jl_b200", "csrc")


class HMParams(C.Structure):
    _fields_ = [("Nx", C.c_int), ("Ny", C.c_int), ("Hx", C.c_int), ("Hy", C.c_int),
                ("south", C.c_double), ("north_poles_latitude", C.c_double),
                ("first_pole_longitude", C.c_double), ("radius", C.c_double),
                ("i0", C.c_int), ("j0", C.c_int), ("nx", C.c_int), ("ny", C.c_int)]


def _build():
    deps = [_SRC] + [os.path.join(_CSRC, n) for n in ("ossg_math.cuh", "ossg_tripolar_core.cuh")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-shared",
                    "-o", _SO, _SRC], check=True, capture_output=True)


def load():
    _build()
    lib = C.CDLL(_SO)
    vpp = C.POINTER(C.c_void_p)
    lib.hm_coords.argtypes = [C.POINTER(HMParams), vpp]
    lib.hm_metrics.argtypes = [C.POINTER(HMParams), vpp, vpp, C.c_int]
    lib.hm_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.hm_haversine.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.hm_quadrilateral.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.hm_halo_source.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)] * 3
    return lib


def _ptrs(arrays):
    arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return C.cast(arr, C.POINTER(C.c_void_p))


def make(size, halo=(4, 4, 4), south=-80, north_poles_latitude=55, first_pole_longitude=70,
         radius=6371.0e3, tile=None):
    Nx, Ny = int(size[0]), int(size[1])
    i0, j0, nx, ny = tile if tile else (1, 1, Nx, Ny)
    return HMParams(Nx, Ny, int(halo[0]), int(halo[1]), float(south), float(north_poles_latitude),
                    float(first_pole_longitude), float(radius), i0, j0, nx, ny)


def coords(lib, p):
    shape = (p.ny + 2 * p.Hy, p.nx + 2 * p.Hx)
    out = [np.empty(shape) for _ in range(8)]
    assert lib.hm_coords(C.byref(p), _ptrs(out)) == 0
    return out


def metrics(lib, p, coords_in, closed_form):
    shape = (p.ny + 2 * p.Hy, p.nx + 2 * p.Hx)
    out = [np.empty(shape) for _ in range(12)]
    cin = [np.ascontiguousarray(c) for c in coords_in]
    assert lib.hm_metrics(C.byref(p), _ptrs(cin), _ptrs(out), int(closed_form)) == 0
    return out


def math(lib, fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    assert lib.hm_math_eval(fn, x.ctypes.data, y.ctypes.data, x.size) == 0
    return y


def haversine(lib, hot, rows):
    """rows: n x 7 (lam_x, hphi1, c1, lam_y, hphi2, c2, R)"""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    y = np.empty(rows.shape[0])
    assert lib.hm_haversine(int(hot), rows.ctypes.data, y.ctypes.data, rows.shape[0]) == 0
    return y


def quadrilateral(lib, hot, rows):
    """rows: n x 8 (lat, lon of corners a, b, c, d in degrees)"""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    y = np.empty(rows.shape[0])
    assert lib.hm_quadrilateral(int(hot), rows.ctypes.data, y.ctypes.data, rows.shape[0]) == 0
    return y
