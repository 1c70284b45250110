"""Small cases for compute-sanitizer (one tool per run):
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
60x30 construction (Float64 + Float32, with_halo), a 4-field fused fill_halo_regions and zipper fill, the
tile-window build of a 2x2 partition, and the emulated 4-rank peer fold (pull, push, bulk push) on one GPU.
Every result is checked against the oracle, so a sanitizer-induced slowdown cannot hide a wrong answer."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import ossg_b200 as ossg  # noqa: E402
from ossg_b200 import _lib  # noqa: E402
from ossg_b200.distributed import local_sizes, tile_of  # noqa: E402
from ossg_b200.grids import build_arrays, make_params  # noqa: E402
from oracle import oracle as orc  # noqa: E402

o = orc.Oracle()
L = _lib.lib()
dev = torch.device("cuda", 0)


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


# ---- construction
ref = o.build(orc.make_params((60, 30, 2)))
g = ossg.TripolarGrid(ossg.GPU(), size=(60, 30, 2))
assert all(same(g.arrays[n].parent.cpu().numpy(), ref[n]) for n in orc.ALL_NAMES)
g32 = ossg.TripolarGrid(ossg.GPU(), ossg.Float32, size=(60, 30, 2))
assert all(same(g32.arrays[n].parent.cpu().numpy(), ref[n].astype(np.float32)) for n in orc.ALL_NAMES)
big = ossg.with_halo((5, 9, 5), g)
refb = o.build(orc.make_params((60, 30, 2), (5, 9, 5)))
assert all(same(big.arrays[n].parent.cpu().numpy(), refb[n]) for n in orc.ALL_NAMES)
for rx in range(2):
    for ry in range(2):
        i0, j0, nx, ny = tile_of(60, 30, 2, 2, rx, ry)
        t = build_arrays(make_params((60, 30, 2), (4, 4, 4), -80, 55, 70, orc.R_EARTH, tile=(i0, j0, nx, ny)), ossg.Float64, dev)
        assert all(same(a.cpu().numpy(), ref[n][j0 - 1:j0 + ny + 7, i0 - 1:i0 + nx + 7]) for n, a in zip(orc.ALL_NAMES, t))
print("construction ok")

# ---- fills
LOCS = [((0, 0), 1), ((1, 0), -1), ((0, 1), -1), ((1, 1), 1)]
Nx, Ny, Nz, H = 60, 30, 3, 4
rng = np.random.default_rng(3)
hosts = [rng.standard_normal((Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)) for _ in LOCS]
for fn, ofn in ((L.ossg_fill_halo_regions_batch, o.fill_halo_regions), (L.ossg_zipper_fill_batch, o.zipper)):
    devs = [torch.as_tensor(h.copy()).cuda() for h in hosts]
    arr = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(d.data_ptr(), 0, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, float(sg)) for d, (loc, sg) in zip(devs, LOCS)])
    _lib.check(fn(arr, 4, None))
    torch.cuda.synchronize()
    for h, d, (loc, sg) in zip(hosts, devs, LOCS):
        assert same(d.cpu().numpy(), ofn(h.copy(), Nx, Ny, Nz, H, H, H, loc[0], loc[1], sg))
print("fills ok")

# ---- emulated 4-rank peer fold (each rank's kernel on its own stream)
# (compute-sanitizer may serialise kernels; ranks that wait for each other then never meet, so the default is
# ONE rank -- the same kernel and memory accesses, folding onto itself; --peer4 runs the 4-rank form)
gNx, Ny, Nz, Hx, Hy, Hz, Rx = 1100, 10, 2, 3, 3, 1, (4 if "--peer4" in sys.argv else 1)
sx = local_sizes(gNx, Rx)
starts = [1 + sum(sx[:r]) for r in range(Rx)]
for push in (0, 1, 2):
    hosts = [rng.standard_normal((Nz + 2 * Hz, Ny + 2 * Hy, gNx + 2 * Hx)) for _ in LOCS]
    tiles = [[torch.as_tensor(np.ascontiguousarray(h[:, :, starts[r] - 1:starts[r] - 1 + sx[r] + 2 * Hx])).cuda() for r in range(Rx)] for h in hosts]
    flags = [torch.zeros(17, dtype=torch.int64, device=dev) for _ in range(Rx)]
    streams = [torch.cuda.Stream() for _ in range(Rx)]
    torch.cuda.synchronize()
    for epoch in (1, 2):
        for r in range(Rx):
            fd = _lib.PeerFoldDesc()
            fd.gNx, fd.me, fd.nranks, fd.epoch, fd.push = gNx, r, Rx, epoch, push
            for q in range(Rx):
                fd.i0[q], fd.nx[q] = starts[q], sx[q]
                fd.flags_peer[q] = None if q == r else flags[q].data_ptr()
            fd.flags_local = flags[r].data_ptr()
            arr = (_lib.PeerField * 4)()
            for k, (loc, sg) in enumerate(LOCS):
                arr[k].f = _lib.FieldDesc(tiles[k][r].data_ptr(), 0, sx[r], Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], 0, 0, float(sg))
                for q in range(Rx):
                    arr[k].peer[q] = tiles[k][q].data_ptr()
            _lib.check(L.ossg_zipper_fill_peer(arr, 4, C.byref(fd), streams[r].cuda_stream))
        torch.cuda.synchronize()
        for k, (loc, sg) in enumerate(LOCS):
            o.zipper(hosts[k], gNx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sg)
            for r in range(Rx):
                assert same(tiles[k][r].cpu().numpy()[:, :, Hx:Hx + sx[r]], hosts[k][:, :, starts[r] - 1 + Hx:starts[r] - 1 + Hx + sx[r]]), (push, k, r, epoch)
print("peer fold ok")
print("sanitize_case done")
