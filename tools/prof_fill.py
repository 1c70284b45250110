"""Workload for ncu captures of the fill kernels (never a bench value), same loops as bench.py:

    --loop8    8 fields per zipper launch (235 MB per fill, more than the L2), 40 fills back to back, no flush
    --zipper4  c,u,v,zeta per zipper launch, 256 MB flush write before every launch
    --fill     the same with ossg_fill_halo_regions_batch (zipper + periodic west/east)

    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,\
lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:zipper_kernel --launch-skip 20 --launch-count 4 \
        --csv --log-file gpurun_out/zipper_loop8.csv python tools/prof_fill.py --loop8
(--cache-control none: the L2 holds what the loop left in it, as in the timed loop.)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import ossg_b200  # noqa: E402,F401
from ossg_b200 import _lib  # noqa: E402

Nx, Ny, Nz, H = 4320, 2160, 100, 4
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
L = _lib.lib()
s = torch.cuda.current_stream(dev)
shape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
nf = 8 if "--loop8" in sys.argv else 4
fields = [torch.zeros(shape, dtype=torch.float64, device=dev) for _ in range(nf)]
descs = (_lib.FieldDesc * nf)(*[_lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, sg)
                                 for t, (loc, sg) in zip(fields, (specs + specs)[:nf])])
if "--loop8" in sys.argv:
    for _ in range(40):
        _lib.check(L.ossg_zipper_fill_batch(descs, 8, s.cuda_stream))
else:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = L.ossg_fill_halo_regions_batch if "--fill" in sys.argv else L.ossg_zipper_fill_batch
    for it in range(8):
        flush.fill_(it)
        _lib.check(fn(descs, 4, s.cuda_stream))
torch.cuda.synchronize()
print("prof_fill done")
