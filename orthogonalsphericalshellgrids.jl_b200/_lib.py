"""Loader of the sm_100a shared library (libossg_b200.so) and the ctypes mirror of
include/ossg_b200.h.  There is no fallback: if the library is missing or no B200-class device
is current, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("OSSG_B200_LIB") or os.path.join(_PKG, "libossg_b200.so")   # (override: kernel experiments)
HEADER_PATH = os.path.join(_ROOT, "include", "ossg_b200.h")
UTIL_LIB_PATH = os.path.join(_PKG, "libossg_b200_testutil.so")
UTIL_HEADER_PATH = os.path.join(_ROOT, "include", "ossg_b200_testutil.h")

# error codes (include/ossg_b200.h)
OSSG_EINVAL, OSSG_EODD_NX, OSSG_EHALO, OSSG_EWINDOW, OSSG_EALIGN, OSSG_ELOCATION, OSSG_EWORKSPACE = -1, -2, -3, -4, -5, -6, -7
OSSG_ENCCL, OSSG_ETIMEOUT = -8, -9
OSSG_CENTER, OSSG_FACE = 0, 1
OSSG_F64, OSSG_F32 = 0, 1
OSSG_LAMBDA1D_F32 = 1

ARRAY_NAMES = ["lam_cc", "lam_fc", "lam_cf", "lam_ff", "phi_cc", "phi_fc", "phi_cf", "phi_ff",
               "dx_cc", "dx_fc", "dx_cf", "dx_ff", "dy_cc", "dy_cf", "dy_fc", "dy_ff",
               "az_cc", "az_fc", "az_cf", "az_ff"]


class TripolarParams(C.Structure):
    _fields_ = [("Nx", C.c_int32), ("Ny", C.c_int32), ("Hx", C.c_int32), ("Hy", C.c_int32),
                ("southernmost_latitude", C.c_double), ("north_poles_latitude", C.c_double),
                ("first_pole_longitude", C.c_double), ("radius", C.c_double),
                ("i0", C.c_int32), ("j0", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class FieldDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("elem_type", C.c_int32),
                ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
                ("Hx", C.c_int32), ("Hy", C.c_int32), ("Hz", C.c_int32),
                ("locx", C.c_int32), ("locy", C.c_int32),
                ("k0", C.c_int32), ("nk", C.c_int32), ("sign", C.c_double)]


class FoldWindow(C.Structure):
    _fields_ = [("gNx", C.c_int32), ("i0", C.c_int32), ("nx", C.c_int32),
                ("src_i0", C.c_int32), ("src_n", C.c_int32)]


OSSG_MAX_PEERS, OSSG_MAX_PEER_BATCH = 8, 16


class ExchangeLayout(C.Structure):
    _fields_ = [("gNx", C.c_int32), ("gNy", C.c_int32), ("Rx", C.c_int32), ("Ry", C.c_int32),
                ("rank", C.c_int32), ("reserved", C.c_int32)]


class FoldSegment(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("field", "src_rank", "dst_rank", "src_i0", "n", "rows", "wrapped", "reserved")]


class IpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class PeerFoldDesc(C.Structure):
    _fields_ = [("gNx", C.c_int32), ("me", C.c_int32), ("nranks", C.c_int32),
                ("i0", C.c_int32 * OSSG_MAX_PEERS), ("nx", C.c_int32 * OSSG_MAX_PEERS),
                ("flags_local", C.c_void_p), ("flags_peer", C.c_void_p * OSSG_MAX_PEERS),
                ("epoch", C.c_uint64), ("push", C.c_int32), ("bulk_waves", C.c_int32)]


class PeerField(C.Structure):
    _fields_ = [("f", FieldDesc), ("peer", C.c_void_p * OSSG_MAX_PEERS)]


class PeerHaloDesc(C.Structure):
    _fields_ = [("me", C.c_int32), ("nranks", C.c_int32),
                ("north", C.c_int32), ("south", C.c_int32), ("east", C.c_int32), ("west", C.c_int32),
                ("phase", C.c_int32), ("reserved", C.c_int32),
                ("flags_local", C.c_void_p), ("flags_peer", C.c_void_p * OSSG_MAX_PEERS),
                ("epoch", C.c_uint64),
                ("stage_to_east", C.c_void_p), ("stage_to_west", C.c_void_p),
                ("stage_from_west", C.c_void_p), ("stage_from_east", C.c_void_p)]


class PeerHaloField(C.Structure):
    _fields_ = [("f", FieldDesc), ("north", C.c_void_p), ("south", C.c_void_p), ("east", C.c_void_p), ("west", C.c_void_p),
                ("ny_north", C.c_int32), ("ny_south", C.c_int32), ("nx_east", C.c_int32), ("nx_west", C.c_int32)]


class OSSGError(RuntimeError):
    """Non-argument failure reported by libossg_b200.so (a CUDA error)."""


_lib = None


def build(verbose: bool = False) -> str:
    """Compile libossg_b200.so for sm_100a in-tree with nvcc (csrc/Makefile)."""
    res = subprocess.run(["make", "-C", os.path.join(_PKG, "csrc")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc build of libossg_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vpp = C.POINTER(C.c_void_p)
    L.ossg_version.restype = C.c_int
    L.ossg_last_error.restype = C.c_char_p
    L.ossg_device_ok.restype = C.c_int
    L.ossg_tripolar_workspace_bytes.argtypes = [C.POINTER(TripolarParams), C.c_int]
    L.ossg_tripolar_workspace_bytes.restype = C.c_size_t
    L.ossg_tripolar_workspace_init.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.ossg_tripolar_table_bytes.argtypes = [C.POINTER(TripolarParams)]
    L.ossg_tripolar_table_bytes.restype = C.c_size_t
    L.ossg_tripolar_build.argtypes = [C.POINTER(TripolarParams), C.c_int, vpp, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ossg_tripolar_build_profiled.argtypes = [C.POINTER(TripolarParams), C.c_int, vpp, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_float)]
    L.ossg_tripolar_coords.argtypes = [C.POINTER(TripolarParams), vpp, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ossg_tripolar_metrics.argtypes = [C.POINTER(TripolarParams), vpp, C.c_int, vpp, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ossg_cartesian_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ossg_zipper_fill.argtypes = [C.POINTER(FieldDesc), C.c_void_p]
    L.ossg_zipper_fill_batch.argtypes = [C.POINTER(FieldDesc), C.c_int, C.c_void_p]
    L.ossg_fill_halo_regions_batch.argtypes = [C.POINTER(FieldDesc), C.c_int, C.c_void_p]
    L.ossg_fold_buffer_elems.argtypes = [C.POINTER(FieldDesc), C.POINTER(FoldWindow)]
    L.ossg_fold_buffer_elems.restype = C.c_size_t
    L.ossg_fold_pack.argtypes = [C.POINTER(FieldDesc), C.POINTER(FoldWindow), C.c_void_p, C.c_void_p]
    L.ossg_fold_unpack.argtypes = [C.POINTER(FieldDesc), C.POINTER(FoldWindow), C.c_void_p, C.c_void_p]
    for name in ("ossg_fold_exchange", "ossg_halo_exchange"):
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_int, C.POINTER(ExchangeLayout), C.c_void_p, C.c_size_t, C.c_void_p]
        getattr(L, name + "_workspace_bytes").argtypes = [C.POINTER(FieldDesc), C.c_int, C.POINTER(ExchangeLayout)]
        getattr(L, name + "_workspace_bytes").restype = C.c_size_t
    L.ossg_fold_exchange_plan.argtypes = [C.POINTER(FieldDesc), C.c_int, C.POINTER(ExchangeLayout), C.POINTER(FoldSegment), C.c_int]
    L.ossg_nccl_get_unique_id.argtypes = [C.c_char_p]
    L.ossg_nccl_comm_init_rank.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_int]
    L.ossg_nccl_comm_destroy.argtypes = [C.c_void_p]
    L.ossg_ipc_export.argtypes = [C.c_void_p, C.POINTER(IpcHandle), C.POINTER(C.c_uint64)]
    L.ossg_ipc_open.argtypes = [C.POINTER(IpcHandle), C.c_uint64, C.POINTER(C.c_void_p)]
    L.ossg_ipc_close.argtypes = [C.POINTER(IpcHandle)]
    L.ossg_zipper_fill_peer.argtypes = [C.POINTER(PeerField), C.c_int, C.POINTER(PeerFoldDesc), C.c_void_p]
    L.ossg_halo_exchange_peer.argtypes = [C.POINTER(PeerHaloField), C.c_int, C.POINTER(PeerHaloDesc), C.c_void_p]
    _lib = L
    return L


_util = None


def testutil() -> C.CDLL:
    """libossg_b200_testutil.so (include/ossg_b200_testutil.h): synthetic fields, probes, self-tests.
    Test and measurement infrastructure -- nothing in the package's product path calls it."""
    global _util
    if _util is not None:
        return _util
    if not os.path.exists(UTIL_LIB_PATH):
        raise ImportError(f"{UTIL_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    U = C.CDLL(UTIL_LIB_PATH)
    U.ossg_testutil_last_error.restype = C.c_char_p
    U.ossg_synth_field.argtypes = [C.POINTER(FieldDesc), C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]
    U.ossg_fp64_peak_probe.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p]
    U.ossg_selftest_sqrt_div.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
    U.ossg_peer_copy_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64,
                                       C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    _util = U
    return U


def check_util(rc: int) -> None:
    if rc == 0:
        return
    msg = testutil().ossg_testutil_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(msg)
    raise OSSGError(f"CUDA error {rc}: {msg}")


def last_error() -> str:
    return lib().ossg_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a return code to the exception the Julia shim would raise: negative -> ArgumentError
    (ValueError here), positive -> CUDA failure."""
    if rc == 0:
        return
    msg = last_error()
    if rc < 0:
        raise ValueError(msg)
    raise OSSGError(f"CUDA error {rc}: {msg}")


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return C.cast(arr, C.POINTER(C.c_void_p))


def declared_symbols(path: str = None) -> list:
    """Function names declared in include/ossg_b200.h (or the given header), for the export test."""
    import re
    text = open(path or HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ossg_[a-z0-9_]+)\s*\(", text)))
