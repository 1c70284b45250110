"""Per-rank construction time of Y-slab windows of the 1/12 degree grid, measured on ONE GPU (the distributed
constructor has no communication: a rank's work is exactly this window).  For the library named by OSSG_B200_LIB
and the frame mode named by OSSG_FRAME.  One line per slab: device time per build, 40 builds back to back, best of 3.
Never a bench value; it exists to choose launch forms for the multi-GPU partitions without an 8-GPU box."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossg_b200 import _lib
from ossg_b200.grids import make_params

L = _lib.lib(); dev = torch.device('cuda', 0)
Nx, Ny, H = 4320, 2160, 4
s = torch.cuda.current_stream(dev)
res = []
TILES = ((1, 1, 0, 0), (1, 2, 0, 1), (1, 4, 0, 3), (1, 8, 0, 0), (1, 8, 0, 3), (1, 8, 0, 7))
if "--xsplit" in sys.argv:
    TILES = ((2, 1, 0, 0), (2, 1, 1, 0), (4, 1, 1, 0), (8, 1, 3, 0), (2, 2, 1, 1), (2, 4, 0, 3), (4, 2, 2, 1))
for (Rx, Ry, rx, r) in TILES:
    ny = Ny // Ry
    nx = Nx // Rx
    p = make_params((Nx, Ny, 100), (H, H, H), -80.0, 55.0, 70.0, 6371.0e3, tile=(rx * nx + 1, r * ny + 1, nx, ny))
    out = [torch.zeros((ny + 2 * H, nx + 2 * H), dtype=torch.float64, device=dev) for _ in range(20)]
    optr = _lib.ptr_array([t.data_ptr() for t in out])
    nb = L.ossg_tripolar_workspace_bytes(C.byref(p), _lib.OSSG_F64)
    ws = torch.zeros(max(nb, 128), dtype=torch.uint8, device=dev)
    _lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), s.cuda_stream))
    for _ in range(5):
        _lib.check(L.ossg_tripolar_build(C.byref(p), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), s.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(s)
        for _ in range(40):
            _lib.check(L.ossg_tripolar_build(C.byref(p), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), s.cuda_stream))
        e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    chk = 0
    for t in out:
        chk ^= int(t.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
    res.append('(%d,%d)#%d,%d %.4f %04x' % (Rx, Ry, rx, r, best, chk & 0xffff))
    del out, ws
print('%-22s frame=%-5s %s' % (os.path.basename(_lib.LIB_PATH), os.environ.get('OSSG_FRAME', 'dflt'), ' | '.join(res)), flush=True)
