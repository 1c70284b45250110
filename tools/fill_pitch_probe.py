"""What bounds the periodic part of fill_halo_regions!?  The same number of rows (2168 x 108 per field, 4 fields: 3.7 M
32-byte halo pieces to read and as many to write) at different row pitches: with a short pitch consecutive rows' pieces
share DRAM rows, with the 1/12 degree pitch (34 KB) every piece opens its own.  Event-timed, 256 MB flush before every
launch, median of 9; the zipper part (proportional to Nx) is timed on its own and subtracted."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossg_b200 import _lib
Ny, Nz, H = 2160, 100, 4
dev = torch.device("cuda", 0); L = _lib.lib(); s = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
def timed(fn, descs):
    ts = []
    for it in range(12):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); _lib.check(fn(descs, 4, s.cuda_stream)); e1.record(s); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print("rows per field %d, halo pieces per launch %.2f M read + %.2f M written" % ((Ny + 2 * H) * (Nz + 2 * H), 8 * (Ny + 2 * H) * (Nz + 2 * H) / 1e6, 8 * (Ny + 2 * H) * (Nz + 2 * H) / 1e6))
for Nx in (64, 128, 256, 512, 1024, 2048, 4320):
    shape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
    fields = [torch.zeros(shape, dtype=torch.float64, device=dev) for _ in specs]
    descs = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, sg)
                                    for t, (loc, sg) in zip(fields, specs)])
    tz = timed(L.ossg_zipper_fill_batch, descs)
    tf = timed(L.ossg_fill_halo_regions_batch, descs)
    print("Nx %5d  row pitch %6d B  field %7.1f MB | zipper %6.1f us  fill_halo_regions! %6.1f us  periodic part ~ %6.1f us" %
          (Nx, (Nx + 2 * H) * 8, shape[0] * shape[1] * shape[2] * 8 / 1e6, tz, tf, tf - tz), flush=True)
    del fields
