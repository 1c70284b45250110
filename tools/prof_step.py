"""Workload for ncu captures (never a bench value): N constructions of the 1/12 degree grid and,
optionally, batched zipper fills of c,u,v,zeta.

    ncu --set full --clock-control none --import-source on -k regex:'metrics_tile|coords_pair' \
        --launch-skip 6 --launch-count 2 -o gpurun_out/prof python tools/prof_step.py [--zipper] [--f32]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import ossg_b200  # noqa: E402,F401
from ossg_b200 import _lib  # noqa: E402
from ossg_b200.grids import make_params  # noqa: E402

Nx, Ny, Nz, H = 4320, 2160, 100, 4
if "--small" in sys.argv:
    Nx, Ny, Nz = 1440, 720, 75
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
L = _lib.lib()
f32 = "--f32" in sys.argv
typ = _lib.OSSG_F32 if f32 else _lib.OSSG_F64
params = make_params((Nx, Ny, Nz), (H, H, H), -80.0, 55.0, 70.0, 6371.0e3)
shape = (Ny + 2 * H, Nx + 2 * H)
out = [torch.empty(shape, dtype=torch.float32 if f32 else torch.float64, device=dev) for _ in range(20)]
optr = _lib.ptr_array([t.data_ptr() for t in out])
nbytes = L.ossg_tripolar_workspace_bytes(C.byref(params), typ)
ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream(dev)
for _ in range(4):
    _lib.check(L.ossg_tripolar_build(C.byref(params), typ, optr, ws.data_ptr(), ws.numel(), s.cuda_stream))
torch.cuda.synchronize()

if "--zipper" in sys.argv:
    fshape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
    specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
    fields = [torch.zeros(fshape, dtype=torch.float64, device=dev) for _ in specs]
    descs = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, sg)
                                    for t, (loc, sg) in zip(fields, specs)])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for it in range(4):
        flush.fill_(it)
        _lib.check(L.ossg_zipper_fill_batch(descs, 4, s.cuda_stream))
        _lib.check(L.ossg_fill_halo_regions_batch(descs, 4, s.cuda_stream))
    torch.cuda.synchronize()
print("prof_step done")
