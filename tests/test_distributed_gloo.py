"""Multi-rank host logic on the CPU: world_size 2 and 4 over gloo (127.0.0.1).  The fold
communication plan, message aggregation and ordinary halo exchange of
ossg_b200.distributed are exercised with a torch-CPU stand-in for the CUDA pack/unpack kernels
(TorchPacker below, written from the buffer layout documented in include/ossg_b200.h); the
result on every rank must equal the same slice of the serial oracle fill.  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TorchPacker:
    """CPU stand-in for ossg_fold_pack / ossg_fold_unpack (same buffer layout:
    [k][row r = 0..Hy-1 (+1 for y-Center)][global column ascending])."""

    def empty(self, f, n):
        return torch.empty(n, dtype=f.data.dtype)

    def stream(self, f):
        return None

    def copy_rows(self, dst, src):
        dst.copy_(src)

    @staticmethod
    def _dims(f):
        g = f.grid
        d = f.descriptor()
        rows = g.Hy + (1 if d.locy == 0 else 0)
        return g, d, rows

    def buffer_elems(self, f, w):
        g, d, rows = self._dims(f)
        return g.Nz * rows * w.src_n

    def pack(self, f, w, buf, stream):
        g, d, rows = self._dims(f)
        b = buf.view(g.Nz, rows, w.src_n)
        for r in range(rows):
            js = (g.Ny - r if d.locy else g.Ny - r - 1) if r < g.Hy else g.Ny
            li0 = w.src_i0 - w.i0 + 1
            b[:, r, :] = f.data[g.Hz:g.Hz + g.Nz, js + g.Hy - 1, li0 + g.Hx - 1:li0 + g.Hx - 1 + w.src_n]

    def unpack(self, f, w, buf, stream):
        g, d, rows = self._dims(f)
        b = buf.view(g.Nz, rows, w.src_n)
        for li in range(1, g.Nx + 1):
            i = w.i0 + li - 1
            s = d.sign
            ip = w.gNx - i + 2 if d.locx else w.gNx - i + 1
            if ip > w.gNx:
                s, ip = abs(s), ip - w.gNx
            c = ip - w.src_i0
            if c < 0 or c >= w.src_n:
                continue
            for r in range(rows):
                if r < g.Hy:
                    jd = g.Ny + r + 1
                elif i > w.gNx // 2:
                    jd = g.Ny
                else:
                    continue
                f.data[g.Hz:g.Hz + g.Nz, jd + g.Hy - 1, li + g.Hx - 1] = s * b[:, r, c]


class StubGrid:
    def __init__(self, arch, size, halo, i0, j0, global_size):
        self.Nx, self.Ny, self.Nz = size
        self.Hx, self.Hy, self.Hz = halo
        self.architecture = arch
        self.dtype = torch.float64
        self.arrays = {}
        self.i0, self.j0 = i0, j0
        self.global_size = global_size
        self.size, self.halo = size, halo


def _worker(rank, world, port, Rx, Ry, gsize, halo, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ossg_b200 as ossg
        from ossg_b200.distributed import distributed_fill_halo_regions, tile_of
        from oracle import oracle as orc
        o = orc.Oracle()
        Nx, Ny, Nz = gsize
        Hx, Hy, Hz = halo
        arch = ossg.Distributed(partition=ossg.Partition(Rx, Ry))
        i0, j0, nx, ny = tile_of(Nx, Ny, Rx, Ry, arch.rx, arch.ry)
        grid = StubGrid(arch, (nx, ny, Nz), halo, i0, j0, gsize)
        C_, F_ = ossg.Center, ossg.Face
        specs = [((C_, C_, C_), (0, 0)), ((F_, C_, C_), (1, 0)), ((C_, F_, C_), (0, 1)), ((F_, F_, C_), (1, 1))]
        fields, wants = [], []
        for k, (loc, code) in enumerate(specs):
            rng = np.random.default_rng(100 + k)            # same global field on every rank
            G = np.full((Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx), np.nan)
            G[Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx] = rng.uniform(-1, 1, (Nz, Ny, Nx))
            local = np.full((Nz + 2 * Hz, ny + 2 * Hy, nx + 2 * Hx), np.nan)
            local[Hz:Hz + Nz, Hy:Hy + ny, Hx:Hx + nx] = G[Hz:Hz + Nz, j0 - 1 + Hy:j0 - 1 + Hy + ny, i0 - 1 + Hx:i0 - 1 + Hx + nx]
            f = ossg.Field(loc, grid, data=torch.from_numpy(local))
            sgn = f.boundary_conditions.north.condition if arch.is_top_row() else ossg.sign(loc[0], loc[1])
            want = o.fill_halo_regions(G.copy(), Nx, Ny, Nz, Hx, Hy, Hz, code[0], code[1], sgn)
            fields.append(f)
            wants.append(want[:, j0 - 1:j0 - 1 + ny + 2 * Hy, i0 - 1:i0 - 1 + nx + 2 * Hx])
        distributed_fill_halo_regions(fields, packer=TorchPacker())
        ok = True
        msg = ""
        for f, want, (loc, code) in zip(fields, wants, specs):
            got = f.data.numpy()
            # z-halo levels are not exchanged by the fold (interior levels only, like the reference's
            # (i, k) kernel range); compare the interior levels, NaN == NaN
            a, b = got[Hz:Hz + Nz], want[Hz:Hz + Nz]
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            if not same.all():
                ok = False
                bad = np.argwhere(~same)[0]
                msg += f"rank {rank} loc {code}: mismatch at {bad.tolist()} got {a[tuple(bad)]} want {b[tuple(bad)]}; "
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("Rx,Ry", [(1, 2), (2, 1), (2, 2), (4, 1)])
def test_distributed_fill_equals_slices_of_the_serial_fill(Rx, Ry):
    world = Rx * Ry
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Rx, Ry, (24, 16, 2), (3, 3, 1), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in sorted(results):
        assert ok, msg
