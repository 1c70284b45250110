"""Per-kernel device times (ossg_tripolar_build_profiled: events between the kernels, frame on its own grid) of the
windows of an 8-rank Y-slab partition of the 1/12 degree grid, on one GPU: which kernel makes the top slab slow?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossg_b200 import _lib
from ossg_b200.grids import make_params
L = _lib.lib(); dev = torch.device('cuda', 0); s = torch.cuda.current_stream(dev)
Nx, Ny, H = 4320, 2160, 4
for r in (0, 3, 6, 7):
    ny = Ny // 8
    p = make_params((Nx, Ny, 100), (H, H, H), -80.0, 55.0, 70.0, 6371.0e3, tile=(1, r * ny + 1, Nx, ny))
    out = [torch.zeros((ny + 2 * H, Nx + 2 * H), dtype=torch.float64, device=dev) for _ in range(20)]
    optr = _lib.ptr_array([t.data_ptr() for t in out])
    nb = L.ossg_tripolar_workspace_bytes(C.byref(p), _lib.OSSG_F64)
    ws = torch.zeros(max(nb, 128), dtype=torch.uint8, device=dev)
    prof = (C.c_float * 4)(); acc = [0.0] * 4; n = 20
    for it in range(n + 5):
        _lib.check(L.ossg_tripolar_build_profiled(C.byref(p), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), s.cuda_stream, prof))
        if it >= 5:
            for k in range(4): acc[k] += prof[k] / n
    print('(1,8) slab %d: tables %.4f coords %.4f tile %.4f frame %.4f ms' % (r, *acc), flush=True)
