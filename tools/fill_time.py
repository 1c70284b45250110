"""Event timing of ossg_fill_halo_regions_batch (c,u,v,zeta, 1/12 degree, Nz = 100, FP64) with a 256 MB flush write before
every launch, median of 9, and a checksum of the fields."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossg_b200 import _lib
Nx, Ny, Nz, H = 4320, 2160, 100, 4
dev = torch.device("cuda", 0); L = _lib.lib(); U = _lib.util_lib() if hasattr(_lib, "util_lib") else None
s = torch.cuda.current_stream(dev)
shape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
g = torch.Generator(device=dev); g.manual_seed(5)
fields = [torch.rand(shape, dtype=torch.float64, device=dev, generator=g) for _ in specs]
descs = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, sg)
                                for t, (loc, sg) in zip(fields, specs)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, nf, clean=False):
    ts = []
    for it in range(12):
        flush.fill_(it)
        if clean:
            flush.sum()                      # write, then read: no dirty lines left in the L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        _lib.check(fn(descs, nf, s.cuda_stream))
        e1.record(s); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
res = [("zipper 4 fields, flushed", timed(L.ossg_zipper_fill_batch, 4)), ("zipper 4 fields, clean flush", timed(L.ossg_zipper_fill_batch, 4, True)),
       ("zipper 1 field, flushed", timed(L.ossg_zipper_fill_batch, 1)), ("fill_halo_regions! 4 fields, flushed", timed(L.ossg_fill_halo_regions_batch, 4))]
chk = 0
for t in fields:
    chk ^= int(t.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
print("%-24s " % os.path.basename(_lib.LIB_PATH) + " | ".join("%s %.1f us" % kv for kv in res) + " | chk %016x" % chk, flush=True)
