"""Multi-GPU check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_distributed_gpu.py [--bench]
One rank per GPU over NCCL.  For every (Rx, Ry) factorisation of N:
  * DistributedTripolarGrid tiles == slices of the serial oracle grid (bit for bit);
  * distributed fill_halo_regions (fold pack -> NCCL send/recv -> unpack, then ordinary halos) of
    c, u, v, zeta == slices of the serial oracle fill (bit for bit).
Two grids: 48x40x3 (tiles as narrow as 6 columns) and 4320x64x4 (tiles of 540 columns at 8 ranks: the
row length of the 1/12 degree benchmark, one bulk-copy chunk per row).  The NCCL arm of the C ABI
(ossg_fold_exchange + ossg_halo_exchange), the eager torch.distributed arm, the peer-memory fold
(pull, push, bulk push) and the peer-memory halo exchange are all checked.
Used by tests/test_gpu_distributed.py (skipped on a single-GPU box) and by hand through gpurun."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ossg_b200 as ossg                                    # noqa: E402
from ossg_b200.distributed import (NcclExchange, PeerFold, PeerHalo, distributed_fill_halo_regions,   # noqa: E402
                                   fold_exchange, ordinary_halo_exchange, tile_of)
from oracle import oracle as orc                            # noqa: E402


def factorisations(n):
    return [(rx, n // rx) for rx in range(1, n + 1) if n % rx == 0]


def check(world, rank, dev, size=(48, 40, 3), halo=(4, 4, 2)):
    o = orc.Oracle()
    Nx, Ny, Nz = size
    Hx, Hy, Hz = halo
    ref = o.build(orc.make_params(size, halo))
    failures = []
    for (Rx, Ry) in factorisations(world):
        if Ny // Ry < Hy + 1 or Nx // Rx < Hx:
            continue                      # (tiles smaller than the halo: rejected by the constructor, tested on the CPU)
        arch = ossg.Distributed(ossg.GPU(dev.index), ossg.Partition(Rx, Ry))
        grid = ossg.TripolarGrid(arch, ossg.Float64, size=size, halo=halo)
        i0, j0, nx, ny = tile_of(Nx, Ny, Rx, Ry, arch.rx, arch.ry)
        assert (grid.Nx, grid.Ny, grid.i0, grid.j0) == (nx, ny, i0, j0)
        for n in orc.ALL_NAMES:
            want = ref[n][j0 - 1:j0 - 1 + ny + 2 * Hy, i0 - 1:i0 - 1 + nx + 2 * Hx]
            got = grid.arrays[n].parent.cpu().numpy()
            if not np.array_equal(got.view(np.uint64), np.ascontiguousarray(want).view(np.uint64)):
                failures.append(f"({Rx},{Ry}) rank {rank}: grid array {n} differs")
        C_, F_ = ossg.Center, ossg.Face
        specs = [((C_, C_, C_), (0, 0)), ((F_, C_, C_), (1, 0)), ((C_, F_, C_), (0, 1)), ((F_, F_, C_), (1, 1))]
        fields, wants = [], []
        for k, (loc, code) in enumerate(specs):
            rng = np.random.default_rng(100 + k)
            G = np.full((Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx), np.nan)
            G[Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx] = rng.uniform(-1, 1, (Nz, Ny, Nx))
            local = np.full((Nz + 2 * Hz, ny + 2 * Hy, nx + 2 * Hx), np.nan)
            local[Hz:Hz + Nz, Hy:Hy + ny, Hx:Hx + nx] = G[Hz:Hz + Nz, j0 - 1 + Hy:j0 - 1 + Hy + ny, i0 - 1 + Hx:i0 - 1 + Hx + nx]
            f = ossg.Field(loc, grid, data=torch.from_numpy(local).to(dev))
            want = o.fill_halo_regions(G.copy(), Nx, Ny, Nz, Hx, Hy, Hz, code[0], code[1], ossg.sign(loc[0], loc[1]))
            fields.append(f)
            wants.append(want[:, j0 - 1:j0 - 1 + ny + 2 * Hy, i0 - 1:i0 - 1 + nx + 2 * Hx])
        pristine = [f.data.clone() for f in fields]
        distributed_fill_halo_regions(fields)
        torch.cuda.synchronize(dev)
        for f, want, (loc, code) in zip(fields, wants, specs):
            a, b = f.data.cpu().numpy()[Hz:Hz + Nz], want[Hz:Hz + Nz]
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            if not same.all():
                bad = np.argwhere(~same)[0]
                failures.append(f"({Rx},{Ry}) rank {rank} loc {code}: fill differs at {bad.tolist()}")
        dist.barrier()
        # the NCCL arm of the C ABI: ossg_fold_exchange + ossg_halo_exchange (pack kernel, one ncclGroup, unpack kernel), two rounds
        nccl = NcclExchange(fields)
        for rep in range(2):
            for f, p0 in zip(fields, pristine):
                f.data.copy_(p0)
            nccl()
            torch.cuda.synchronize(dev)
            for f, want, (loc, code) in zip(fields, wants, specs):
                a, b = f.data.cpu().numpy()[Hz:Hz + Nz], want[Hz:Hz + Nz]
                same = (a == b) | (np.isnan(a) & np.isnan(b))
                if not same.all():
                    bad = np.argwhere(~same)[0]
                    failures.append(f"({Rx},{Ry}) rank {rank} loc {code}: C-ABI NCCL fill (round {rep + 1}) differs at {bad.tolist()}")
            dist.barrier()
        # the same fill with the fold served by the one-kernel NVLink peer path (pull, push and
        # bulk-push forms), three epochs each
        for push in (False, True, "bulk"):
            for f, p0 in zip(fields, pristine):
                f.data.copy_(p0)
            torch.cuda.synchronize(dev)
            dist.barrier()
            peer = PeerFold(fields, push=push)
            for rep in range(3):
                if rep:
                    for f, p0 in zip(fields, pristine):
                        f.data.copy_(p0)
                peer()
                ordinary_halo_exchange(fields)
                torch.cuda.synchronize(dev)
                for f, want, (loc, code) in zip(fields, wants, specs):
                    a, b = f.data.cpu().numpy()[Hz:Hz + Nz], want[Hz:Hz + Nz]
                    same = (a == b) | (np.isnan(a) & np.isnan(b))
                    if not same.all():
                        bad = np.argwhere(~same)[0]
                        failures.append(f"({Rx},{Ry}) rank {rank} loc {code}: PEER fill push={push} (epoch {rep + 1}) "
                                        f"differs at {bad.tolist()}")
                if peer.active and peer.timed_out():
                    failures.append(f"({Rx},{Ry}) rank {rank}: peer kernel timed out waiting for a partner")
                dist.barrier()
        # the whole fill (fold + south/north + west/east) over NVLink peer memory, two epochs
        for f, p0 in zip(fields, pristine):
            f.data.copy_(p0)
        torch.cuda.synchronize(dev)
        dist.barrier()
        phalo = PeerHalo(fields)
        for rep in range(2):
            if rep:
                for f, p0 in zip(fields, pristine):
                    f.data.copy_(p0)
            phalo()
            torch.cuda.synchronize(dev)
            for f, want, (loc, code) in zip(fields, wants, specs):
                a, b = f.data.cpu().numpy()[Hz:Hz + Nz], want[Hz:Hz + Nz]
                same = (a == b) | (np.isnan(a) & np.isnan(b))
                if not same.all():
                    bad = np.argwhere(~same)[0]
                    failures.append(f"({Rx},{Ry}) rank {rank} loc {code}: PeerHalo fill (epoch {rep + 1}) differs at {bad.tolist()}")
            if phalo.timed_out():
                failures.append(f"({Rx},{Ry}) rank {rank}: PeerHalo kernel timed out waiting for a partner")
            dist.barrier()
        # Float32 fields through the same kernels (4-byte elements: the 16-byte vector pieces and their
        # unaligned scalar fallback), against the FP64 reference rounded once (the fill is a copy with a sign)
        f32 = [ossg.Field(f.loc, grid, data=p0.to(torch.float32)) for f, p0 in zip(fields, pristine)]
        torch.cuda.synchronize(dev)
        dist.barrier()
        phalo32 = PeerHalo(f32)
        phalo32()
        torch.cuda.synchronize(dev)
        for f, want, (loc, code) in zip(f32, wants, specs):
            a, b = f.data.cpu().numpy()[Hz:Hz + Nz], want[Hz:Hz + Nz].astype(np.float32)
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            if not same.all():
                bad = np.argwhere(~same)[0]
                failures.append(f"({Rx},{Ry}) rank {rank} loc {code}: Float32 PeerHalo fill differs at {bad.tolist()}")
        dist.barrier()
    return failures


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small-only", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = check(world, rank, dev)
    if not args.small_only:
        failures += [f"[4320x64x4] {x}" for x in check(world, rank, dev, size=(4320, 64, 4), halo=(4, 4, 2))]
    gathered = [None] * world
    dist.all_gather_object(gathered, failures)
    out = {"world": world, "failures": [x for fl in gathered for x in fl]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if out["failures"] else 0)


if __name__ == "__main__":
    main()
