"""Device timing + output checksum of the 1/12 degree construction for the library named by
OSSG_B200_LIB (kernel experiments: tile shapes etc.).  Prints one line; no parity judgement here,
the checksum is compared across variants (same arithmetic -> same bits)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossg_b200 import _lib
from ossg_b200.grids import make_params

L = _lib.lib(); dev = torch.device('cuda', 0)
Nx, Ny, H = 4320, 2160, 4
p = make_params((Nx, Ny, 100), (H, H, H), -80.0, 55.0, 70.0, 6371.0e3)
out = [torch.zeros((Ny + 2 * H, Nx + 2 * H), dtype=torch.float64, device=dev) for _ in range(20)]
optr = _lib.ptr_array([t.data_ptr() for t in out])
nb = L.ossg_tripolar_workspace_bytes(C.byref(p), _lib.OSSG_F64)
ws = torch.zeros(max(nb, 128), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream(dev)
prof = (C.c_float * 4)(); acc = [0.0] * 4; n = 20
for it in range(n + 5):
    _lib.check(L.ossg_tripolar_build_profiled(C.byref(p), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), s.cuda_stream, prof))
    if it >= 5:
        for k in range(4):
            acc[k] += prof[k] / n
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record(s)
    for _ in range(20):
        _lib.check(L.ossg_tripolar_build(C.byref(p), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), s.cuda_stream))
    e1.record(s); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
chk = 0
for t in out:
    chk ^= int(t.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
print('%-40s tables %.4f coords %.4f tile %.4f frame %.4f | step %.4f ms | chk %016x' %
      (os.path.basename(_lib.LIB_PATH), *acc, best, chk), flush=True)
