#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 tripolar-grid path.

    python bench.py --gpus N --steps K --warmup W            (own arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement of the reference)

Metric (BASELINE.json): grid cells/s for coords+metrics (FP64); zipper halo-fill GB/s.
A "step" is one full construction of the 20 padded arrays of the 1/12 degree TripolarGrid
(4320x2160x100, halo 4) -- the configuration the north-star target is quoted on, which fits one
GPU.  At N > 1 the same global grid is partitioned (x, y) = (Rx, Ry) and every rank builds its
own window (no data-path collective; strong scaling of the named configuration).

One JSON line is printed by rank 0.  See DESIGN.md section "Measurement" for every key.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(Nx=4320, Ny=2160, Nz=100, halo=4, south=-80.0, north_poles_latitude=55.0,
                first_pole_longitude=70.0, radius=6371.0e3)
CPU_SAMPLE = dict(Nx=1440, Ny=720)          # fallback CPU sample: 1/9 of the workload's cells per step
CPU_BUDGET_S = 150.0                        # the CPU arm runs the FULL workload unless that would exceed this
METRIC = "grid cells/s for coords+metrics (FP64)"
UNIT = "cells/s"

# FP64 pipe work per launch, counted by ncu on the 1/12 degree workload (SASS opcode mix of the
# committed capture profiles/r02ac_*_opcodes.md, taken at the round-2 kernels of HEAD: DFMA + DMUL + DADD + DSETP
# thread-instructions):
#   metrics_tile_kernel : 843.7 per tile cell (4320 x 2158 cells)        245.79 M warp-instructions for the tiles
#                         (the capture's 251.59 M include the halo frame riding in this grid: 5.80 M, counted on its own
#                         grid in capture r02d).  Round 1: 915.7 per cell; the difference is FP64 COMPARISONS that became
#                         integer max/min and integer comparisons of high words -- the arithmetic is unchanged, so the
#                         FP64 fractions below fall while the kernel gets faster.
#   coords_pair_kernel  : 194.9 per padded element (4328 x 2168 elements)  57.16 M warp-instructions
# and the DRAM traffic of one metrics_tile_kernel launch from the same capture
# (dram__bytes_read.sum + dram__bytes_write.sum = 616.0 MB + 888.7 MB; algorithmic 1491.6 MB).
# These are properties of the code, not of the timing; refresh them whenever the kernels change.
FP64_INST_PER_CELL_TILE = 843.7
FP64_INST_FRAME = 5796220 * 32          # thread-instructions of the whole-grid halo frame (scaled by the frame size of a window below)
FP64_INST_PER_ELEM_COORDS = 194.9
# The same without the comparisons (DFMA + DMUL + DADD only), whole build incl. the frame: 249.61 M (tile + frame) + 54.36 M
# (coordinates) warp-instructions in capture r02ac; capture r02d (start of round 2): 249.22 M + ~5.3 M (frame, own grid) + 54.36 M.
FP64_ARITH_WARP_INST_STEP = 249613243 + 54357486
NCU_TRAFFIC_TILE_BYTES = 616029440 + 888712192
PROFILE_TAG = "r02ac"


def partition_for(n):
    """Headline construction partition at N GPUs: Y slabs, the reference's own (and only) distributed
    layout (src/distributed_tripolar_grid.jl:28-31).  Every rank keeps the full x extent, so the
    coordinate kernel's mirror-pair form applies.  Every other factorisation of N is timed next to it
    (`partition_sweep`) and checked against the oracle (`parity`)."""
    return (1, n)


def factorisations(n):
    return [(rx, n // rx) for rx in range(1, n + 1) if n % rx == 0]


def mixed_partition(n):
    """the (x, y) partition of the distributed fill benchmark / parity: both directions split when N allows"""
    return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(n, (n, 1))


# ------------------------------------------------------------------ independent host restatement of the synthetic field
def host_synth_levels(seed, Nx, Ny, H, ks):
    """numpy restatement of libossg_b200_testutil's ossg_synth_field (csrc/ossg_util.cu hash_value) for the GLOBAL
    field at the 1-based interior levels `ks`: parent array (len(ks), Ny+2H, Nx+2H), halos poisoned with
    the NaN payload 0x7ff8dead0000beef.  Independent of the device code it checks."""
    import numpy as np
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def splitmix(x):
        with np.errstate(over="ignore"):
            x = x + np.uint64(0x9E3779B97F4A7C15)
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return (x ^ (x >> np.uint64(31))) & M

    out = np.empty((len(ks), Ny + 2 * H, Nx + 2 * H), dtype=np.float64)
    out.view(np.uint64)[...] = np.uint64(0x7ff8dead0000beef)
    gi = np.arange(1, Nx + 1, dtype=np.uint64)[None, :]
    gj = np.arange(1, Ny + 1, dtype=np.uint64)[:, None]
    for n, k in enumerate(ks):
        key = gi | (gj << np.uint64(21)) | (np.uint64(k) << np.uint64(42))
        h = splitmix(np.uint64(seed) ^ splitmix(key))
        out[n, H:H + Ny, H:H + Nx] = (h >> np.uint64(40)).astype(np.int64).astype(np.float64) * (1.0 / 8388608.0) - 1.0
    return out


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm (oracle = port of the reference)
def cpu_construct_rate(threads, steps, warmup):
    """cells/s of the CPU restatement (glibc libm, literal per-point evaluation like the
    reference's kernels, row-parallel over `threads`).  The sample is the FULL workload (the 1/12
    degree grid, ~1-5 s per step on 8-24 host threads) as long as warmup + steps of it fit
    CPU_BUDGET_S -- projected from the first full-size step -- otherwise the 1440x720 grid (1/9 of
    the cells).  Returns (cells/s, seconds per step, description of the sample)."""
    from oracle import oracle as orc
    o = orc.Oracle(libm=True)
    o.set_threads(threads)

    def params(Nx, Ny):
        return orc.make_params((Nx, Ny, 1), (WORKLOAD["halo"],) * 3, WORKLOAD["south"], WORKLOAD["north_poles_latitude"],
                               WORKLOAD["first_pole_longitude"], WORKLOAD["radius"])

    Nx, Ny = WORKLOAD["Nx"], WORKLOAD["Ny"]
    p = params(Nx, Ny)
    t0 = time.perf_counter()
    o.build(p)                                   # first full-size step: warm-up and the projection
    first = time.perf_counter() - t0
    what = f"the full workload ({Nx}x{Ny} grid) per step"
    if first * (steps + max(warmup, 1)) > CPU_BUDGET_S:
        Nx, Ny = CPU_SAMPLE["Nx"], CPU_SAMPLE["Ny"]
        p = params(Nx, Ny)
        what = f"{Nx}x{Ny} grid (1/9 of the workload's cells) per step"
        o.build(p)
    for _ in range(max(warmup, 1) - 1):
        o.build(p)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.build(p)
    dt = (time.perf_counter() - t0) / steps
    return Nx * Ny / dt, dt, what


def cpu_fill_rates():
    """The reference's fill_halo_regions! on the host (north_star: "next to it goes the reference's CPU
    construction and fill_halo_regions!"): the oracle's zipper and zipper + periodic-x fill of c,u,v,zeta
    on the 1/4 degree fields (1440x720x75, halo 4, FP64; a bounded sample of the 1/12 degree fields, which
    are 8.1 GB each), one thread -- the oracle is a scalar port of one KA work-item loop."""
    import numpy as np
    from oracle import oracle as orc
    o = orc.Oracle()
    Nx, Ny, Nz, H = 1440, 720, 75, 4
    specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
    fields = [np.full((Nz + 2 * H, Ny + 2 * H, Nx + 2 * H), 0.5 + k) for k in range(len(specs))]   # (a fill is a copy: values do not matter)
    algo = sum(2 * H * Nx * Nz * 8 + (Nx * Nz * 8 if loc[1] == 0 else 0) for loc, _ in specs)
    rows = (Ny + 2 * H) * (Nz + 2 * H)
    algo_full = algo + len(specs) * rows * 2 * (2 * H) * 8
    out = {}
    for name, fn, b in (("zipper_batch4", o.zipper, algo), ("fill_halo_regions_batch4", o.fill_halo_regions, algo_full)):
        reps = 3
        for it in range(reps + 1):
            if it == 1:
                t0 = time.perf_counter()
            for c, (loc, s) in zip(fields, specs):
                fn(c, Nx, Ny, Nz, H, H, H, loc[0], loc[1], s)
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"ms": dt * 1e3, "algorithmic_bytes": b, "gbs": b / dt / 1e9}
    out.update({"cores": 1, "kind": "port",
                "sample": f"{Nx}x{Ny}x{Nz} FP64 fields c,u,v,zeta (1/4 degree), CPU restatement of reference v0.2.1"})
    # the same on the 1/12 degree horizontal size with the four fields filled concurrently (one host thread per field;
    # the fill is a per-field loop in the reference too): a bounded sample of the levels, the whole fields are 8.1 GB each
    import threading
    Nx, Ny, Nz = WORKLOAD["Nx"], WORKLOAD["Ny"], 10
    fields = [np.full((Nz + 2 * H, Ny + 2 * H, Nx + 2 * H), 0.5 + k) for k in range(len(specs))]
    algo = sum(2 * H * Nx * Nz * 8 + (Nx * Nz * 8 if loc[1] == 0 else 0) for loc, _ in specs)
    algo_full = algo + len(specs) * (Ny + 2 * H) * (Nz + 2 * H) * 2 * (2 * H) * 8
    par = {}
    for name, fn, b in (("zipper_batch4", o.zipper, algo), ("fill_halo_regions_batch4", o.fill_halo_regions, algo_full)):
        def work(c, loc, sg):
            fn(c, Nx, Ny, Nz, H, H, H, loc[0], loc[1], sg)
        reps = 3
        for it in range(reps + 1):
            if it == 1:
                t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(c, loc, sg)) for c, (loc, sg) in zip(fields, specs)]
            [t.start() for t in th]
            [t.join() for t in th]
        dt = (time.perf_counter() - t0) / reps
        par[name] = {"ms": dt * 1e3, "algorithmic_bytes": b, "gbs": b / dt / 1e9}
    par.update({"cores": len(specs), "kind": "port",
                "sample": f"{Nx}x{Ny} (1/12 degree) x {Nz} of the 100 levels, FP64 fields c,u,v,zeta filled concurrently, one host thread per field"})
    out["four_threads_twelfth_degree"] = par
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_all = time.perf_counter()
    rate, dt, what = cpu_construct_rate(threads, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": what + ", CPU restatement of reference v0.2.1 with glibc libm; Julia is not installed"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    Rx, Ry = partition_for(n)
    return {"workload": f"TripolarGrid {WORKLOAD['Nx']}x{WORKLOAD['Ny']}x{WORKLOAD['Nz']} (1/12 degree), halo 4, FP64, "
                        "full construction: 8 coordinate + 12 metric padded arrays",
            "partition": f"(x,y)=({Rx},{Ry})", "l2": "outputs 1.5 GB per step, larger than the 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------ own arm
class Ctx:
    """what every section needs: libraries, device, stream, rank layout"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from ossg_b200 import _lib
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa_node = pin_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on stdout when the first communicator is created; stdout must
            # carry the one JSON line only, so fd 1 points at stderr until that has happened
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                import datetime
                dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
                dist.barrier()
                torch.cuda.synchronize(self.dev)
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.L = _lib.lib()
        self.U = _lib.testutil()
        if self.L.ossg_device_ok() != 0:
            raise SystemExit("no usable CUDA device: " + _lib.last_error())
        self.stream = torch.cuda.current_stream(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([int(x)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def timed(self, fn, reps, warm=3):
        """device ms per call of fn: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(reps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / reps)


def pin_to_gpu_numa_node(index):
    """run this process (and place its pinned host buffers, first touch) on the NUMA node the GPU hangs off;
    best effort, silent when sysfs does not say"""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                               capture_output=True, text=True, timeout=10).stdout.strip()
            bus = q[-12:].lower() if q else None
        if not bus:
            return None
        node = int(open(f"/sys/bus/pci/devices/{str(bus).lower()}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:               # noqa: BLE001
        return None


class Construction:
    """one rank's window of the workload grid under an (Rx, Ry) partition: contiguous output block,
    own workspace (header initialised once, so steady-state builds find their 1-D tables in place)"""

    def __init__(self, ctx, Rx, Ry):
        from ossg_b200.distributed import tile_of
        from ossg_b200.grids import make_params
        torch, _lib, L, W = ctx.torch, ctx._lib, ctx.L, WORKLOAD
        self.ctx, self.Rx, self.Ry = ctx, Rx, Ry
        rx, ry = ctx.rank // Ry, ctx.rank % Ry
        self.tile = tile_of(W["Nx"], W["Ny"], Rx, Ry, rx, ry)
        i0, j0, nx, ny = self.tile
        H = W["halo"]
        self.params = make_params((W["Nx"], W["Ny"], W["Nz"]), (H, H, H), W["south"], W["north_poles_latitude"],
                                  W["first_pole_longitude"], W["radius"], tile=self.tile)
        self.shape = (ny + 2 * H, nx + 2 * H)
        self.block = torch.empty((20,) + self.shape, dtype=torch.float64, device=ctx.dev)
        self.optr = _lib.ptr_array([self.block[k].data_ptr() for k in range(20)])
        nbytes = L.ossg_tripolar_workspace_bytes(C.byref(self.params), _lib.OSSG_F64)
        self.ws = torch.empty(max(nbytes, 128), dtype=torch.uint8, device=ctx.dev)
        _lib.check(L.ossg_tripolar_workspace_init(self.ws.data_ptr(), self.ws.numel(), ctx.stream.cuda_stream))

    def step(self):
        c = self.ctx
        c._lib.check(c.L.ossg_tripolar_build(C.byref(self.params), c._lib.OSSG_F64, self.optr, self.ws.data_ptr(), self.ws.numel(),
                                             c.stream.cuda_stream))

    def cold_step(self):
        """a build that cannot reuse the tables (fresh header): what the first constructor call of a run costs"""
        c = self.ctx
        c._lib.check(c.L.ossg_tripolar_workspace_init(self.ws.data_ptr(), self.ws.numel(), c.stream.cuda_stream))
        self.step()


def oracle_reference_to_shm(ctx):
    """rank 0 builds the workload grid with the deterministic CPU oracle (all host threads) and leaves the 20
    arrays in /dev/shm; every rank maps them.  Returns ({name: memmap}, seconds, cleanup)."""
    import numpy as np
    from oracle import oracle as orc
    W = WORKLOAD
    d = f"/dev/shm/ossg_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if ctx.world > 1 else os.getpid()}"
    secs = 0.0
    if ctx.rank == 0:
        os.makedirs(d, exist_ok=True)
        o = orc.Oracle()
        o.set_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        ref = o.build(orc.make_params((W["Nx"], W["Ny"], 1), (W["halo"],) * 3, W["south"], W["north_poles_latitude"],
                                      W["first_pole_longitude"], W["radius"]))
        secs = time.perf_counter() - t0
        for n in orc.ALL_NAMES:
            np.save(os.path.join(d, n + ".npy"), ref[n])
        del ref
    if ctx.world > 1:
        box = [d]
        ctx.dist.broadcast_object_list(box, src=0)
        d = box[0]
    ctx.barrier()
    ref = {n: np.load(os.path.join(d, n + ".npy"), mmap_mode="r") for n in orc.ALL_NAMES}

    def cleanup():
        ctx.barrier()
        if ctx.rank == 0:
            for n in orc.ALL_NAMES:
                try:
                    os.remove(os.path.join(d, n + ".npy"))
                except OSError:
                    pass
            try:
                os.rmdir(d)
            except OSError:
                pass
    return ref, secs, cleanup


def parity_construction(ctx, partitions):
    """every rank's 20 arrays under each partition against the matching slice of the oracle's serial grid
    (src/distributed_tripolar_grid.jl:112-120: a tile IS that slice), bit for bit"""
    import numpy as np
    from oracle import oracle as orc
    ref, secs, cleanup = oracle_reference_to_shm(ctx)
    H = WORKLOAD["halo"]
    out = {"oracle": "deterministic CPU restatement, full 1/12 degree grid, built once on rank 0", "oracle_build_s": secs}
    try:
        for (Rx, Ry) in partitions:
            con = Construction(ctx, Rx, Ry)
            con.step()
            ctx.torch.cuda.synchronize(ctx.dev)
            i0, j0, nx, ny = con.tile
            bad, worst = 0, 0.0
            for k, n in enumerate(orc.ALL_NAMES):
                got = con.block[k].cpu().numpy()
                want = np.ascontiguousarray(ref[n][j0 - 1:j0 - 1 + ny + 2 * H, i0 - 1:i0 - 1 + nx + 2 * H])
                diff = got.view(np.uint64) != want.view(np.uint64)
                nb = int(diff.sum())
                if nb:
                    bad += nb
                    denom = np.maximum(np.abs(want[diff]), np.finfo(np.float64).tiny)
                    worst = max(worst, float((np.abs(got[diff] - want[diff]) / denom).max()))
            total = ctx.sum_over_ranks(bad)
            out[f"({Rx},{Ry})"] = {"mismatches": total, "max_rel_err": ctx.max_over_ranks(worst), "arrays": 20,
                                   "elements_per_array_global": (WORKLOAD["Nx"] + 2 * H) * (WORKLOAD["Ny"] + 2 * H)}
            del con
            ctx.torch.cuda.empty_cache()
    finally:
        cleanup()
    out["mismatches"] = sum(v["mismatches"] for v in out.values() if isinstance(v, dict))
    return out


def run_own(args):
    # a partner that never shows up must end this run, not stall it for minutes (library default: 120 s)
    os.environ.setdefault("OSSG_PEER_TIMEOUT_MS", "30000")
    ctx = Ctx(args)
    torch, dist, _lib, L = ctx.torch, ctx.dist, ctx._lib, ctx.L
    import ossg_b200 as ossg
    world, rank, dev, stream = ctx.world, ctx.rank, ctx.dev, ctx.stream
    W = WORKLOAD
    H = W["halo"]
    Rx, Ry = partition_for(world)
    con = Construction(ctx, Rx, Ry)
    i0, j0, nx, ny = con.tile
    shape = con.shape

    # ---- the timed region: K steps of the full construction (tables in place after the first step, as in any
    # run that constructs more than one grid of a size: with_halo, reconstruct_global_grid)
    for _ in range(max(3, args.warmup)):
        con.step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record(stream)
    for _ in range(args.steps):
        con.step()
    e1.record(stream)
    ctx.barrier()
    ms_step = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    cells_global = W["Nx"] * W["Ny"]
    value = cells_global / (ms_step * 1e-3)
    # the same with a fresh workspace header before every build (the 1-D tables recomputed: a first constructor call)
    ms_cold = ctx.timed(con.cold_step, max(3, min(args.steps, 10)))

    # ---- per-kernel device times (events between the kernels), same workload
    prof = (C.c_float * 4)()
    acc = [0.0] * 4
    nprof = max(3, min(args.steps, 10))
    # (one untimed call first: the stand-alone frame kernel of the profiled build is not part of the ordinary build,
    # its first launch pays CUDA's lazy module loading)
    _lib.check(L.ossg_tripolar_build_profiled(C.byref(con.params), _lib.OSSG_F64, con.optr, con.ws.data_ptr(), con.ws.numel(),
                                              stream.cuda_stream, prof))
    for _ in range(nprof):
        _lib.check(L.ossg_tripolar_build_profiled(C.byref(con.params), _lib.OSSG_F64, con.optr, con.ws.data_ptr(), con.ws.numel(),
                                                  stream.cuda_stream, prof))
        for k in range(4):
            acc[k] += prof[k] / nprof
    kernel_ms = dict(zip(["tables", "coords", "metrics_tile", "metrics_frame"], acc))

    # ---- end to end with host buffers: construct + ONE device-to-host copy of the contiguous 20-array block
    # into pinned memory allocated on the GPU's NUMA node (what TripolarGrid(CPU()) delivers)
    host = torch.empty((20,) + shape, dtype=torch.float64, pin_memory=True)
    pin_params = bytes(con.params)       # the only input: 64 bytes of scalars, passed as kernel arguments

    def e2e_step():
        con.step()
        host.copy_(con.block, non_blocking=True)

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_local = (time.perf_counter() - t0) / e2e_steps
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(e2e_local)
    d2h_bytes = 20 * shape[0] * shape[1] * 8
    clocks = sampler.stop() if rank == 0 else None
    del host

    # ---- wall clock of the public constructor call (allocation of the 20 arrays and the workspace included)
    def public_call():
        if world == 1:
            g = ossg.TripolarGrid(ossg.GPU(dev.index), size=(W["Nx"], W["Ny"], W["Nz"]), halo=(H, H, H))
        else:
            g = ossg.TripolarGrid(ossg.Distributed(ossg.GPU(dev.index), ossg.Partition(Rx, Ry)), size=(W["Nx"], W["Ny"], W["Nz"]), halo=(H, H, H))
        torch.cuda.synchronize(dev)
        return g
    del con.block
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    public_call()                       # first call: the 20 output arrays come from cudaMalloc
    public_first_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
    public_call()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        public_call()                   # steady state: the arrays of the grid dropped before are reused by torch's allocator
    public_ms = ctx.max_over_ranks((time.perf_counter() - t0) / 5 * 1e3)
    torch.cuda.empty_cache()

    # ---- FP64 peak probe (roofline denominator for the FP64-bound kernels)
    ms = C.c_float(); fl = C.c_double()
    _lib.check_util(ctx.U.ossg_fp64_peak_probe(20000, C.byref(ms), C.byref(fl), stream.cuda_stream))
    fp64_peak_tflops = fl.value / (ms.value * 1e-3) / 1e12
    del con
    torch.cuda.empty_cache()

    # ---- parity under the driver's eyes: this rank's arrays against the oracle, for the headline partition, (N,1)
    # and one mixed factorisation (N = 1: the serial grid)
    parts = [(Rx, Ry)]
    for cand in ((world, 1), mixed_partition(world)):
        if cand not in parts:
            parts.append(cand)
    parity = {"construction": parity_construction(ctx, parts)}

    extra = {}
    if world == 1:
        zipper = bench_zipper(ctx, L, _lib, torch, dev, stream, W, nx, ny)
        par = zipper.pop("parity", None) if isinstance(zipper, dict) else None
        if par is not None:
            parity["fill_halo_regions"] = par
        extra["configs"] = bench_small_configs(ctx)
    else:
        # every factorisation of N, timed like the headline (BASELINE config 4)
        sweep = {}
        for (sx_, sy_) in factorisations(world):
            c2 = Construction(ctx, sx_, sy_)
            t = ctx.timed(c2.step, max(5, min(args.steps, 20)))
            sweep[f"({sx_},{sy_})"] = {"ms_per_step": t, "cells_per_s": cells_global / (t * 1e-3),
                                       "coords_kernel": "mirror pairs" if sx_ == 1 else "per element"}
            del c2
            torch.cuda.empty_cache()
        extra["partition_sweep"] = sweep
        # (a failure of the peer-memory plumbing -- e.g. CUDA IPC unavailable in the container -- must not
        # take the construction number down with it)
        zipper = {}
        for key, fn in (("fold_exchange", bench_fold_exchange), ("distributed_fill_halo_regions", bench_distributed_fill)):
            try:
                zipper[key] = fn(ctx)
                par = zipper[key].pop("parity", None)
                if par is not None:
                    parity[key] = par
            except Exception as e:          # noqa: BLE001
                zipper[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
                parity[key] = {"mismatches": 1, "error": "section failed before its parity check"}
            try:
                torch.cuda.empty_cache()
            except Exception:               # noqa: BLE001  (a dead context after a trapped kernel: the line below still goes out)
                pass
    parity["mismatches"] = sum(int(v.get("mismatches", 0)) for v in parity.values() if isinstance(v, dict))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if parity["mismatches"]:
            sys.exit(3)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tile_cells = nx * (ny - 2)                      # cells the tile kernel covers (rows 2..Ny-1)
    # algorithmic bytes of the dominant kernel per cell: reads 8 coordinate values, writes 12 metrics
    tile_bytes = tile_cells * (8 + 12) * 8
    tile_ms = kernel_ms["metrics_tile"]
    achieved = tile_bytes / (tile_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "metrics_tile_kernel<double>", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": NCU_TRAFFIC_TILE_BYTES if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": tile_bytes,
                "note": "this kernel is FP64-pipe bound, not HBM bound; see roofline_fp64"}
    roofline_fp64 = {"bound": "fp64_pipe", "peak": fp64_peak_tflops, "unit": "TFLOP/s (DFMA = 2 flop)",
                     "peak_source": "ossg_fp64_peak_probe on this device, same run",
                     "tile_ns_per_cell": tile_ms * 1e6 / tile_cells}
    if FP64_INST_PER_CELL_TILE:
        peak_inst = fp64_peak_tflops * 1e12 / 2            # FP64 thread-instructions per second
        inst_tile = FP64_INST_PER_CELL_TILE * tile_cells
        inst_coords = FP64_INST_PER_ELEM_COORDS * shape[0] * shape[1]
        roofline_fp64.update({
            "kernel": "metrics_tile_kernel<double>",
            "achieved": 2 * inst_tile / (tile_ms * 1e-3) / 1e12,       # DFMA-equivalent TFLOP/s
            "frac": inst_tile / (tile_ms * 1e-3) / peak_inst,
            "coords_kernel_frac": inst_coords / (kernel_ms["coords"] * 1e-3) / peak_inst,
            "whole_step_frac": (inst_tile + inst_coords + FP64_INST_FRAME / world) / (ms_step * 1e-3) / peak_inst,
            # arithmetic only (no DSETP): comparable across the round -- the range tests that were FP64 comparisons in round 1
            # are integer instructions now, which lowers the fractions above while the kernels get faster
            "whole_step_arith_frac": (FP64_ARITH_WARP_INST_STEP * 32 / world) / (ms_step * 1e-3) / peak_inst if world == 1 else None,
            "fp64_inst_source": f"ncu SASS opcode counts, profiles/{PROFILE_TAG}_*_opcodes.md"})

    roofline_zipper = None
    if world == 1 and isinstance(zipper, dict):
        roofline_zipper = zipper_roofline(zipper, hbm_peak, peak_src)
    cpu_threads = os.cpu_count() or 1
    cpu_rate, cpu_dt, cpu_what = cpu_construct_rate(cpu_threads, 3, 1)
    if world == 1 and isinstance(zipper, dict):
        zipper["cpu_baseline"] = cpu_fill_rates()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "kernel_ms": kernel_ms,
        "first_call": {"ms_per_step": ms_cold, "cells_per_s": cells_global / (ms_cold * 1e-3),
                       "what": "workspace header reset before every build: the 1-D tables are recomputed (tables_kernel's full cost)"},
        "roofline": roofline, "roofline_fp64": roofline_fp64, "roofline_zipper": roofline_zipper,
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                         "sample": cpu_what + ", 3 steps, CPU restatement of reference v0.2.1 (glibc libm); "
                                   "Julia is not installed"},
        "e2e": {"value": cells_global / e2e_s, "unit": UNIT, "h2d_bytes_per_step": len(pin_params),
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s * 1e3,
                "what": "ossg_tripolar_build + ONE copy of the contiguous 20-array block to a pinned host buffer on the GPU's NUMA node "
                        "(TripolarGrid(CPU()) equivalent)",
                "numa_node_of_rank0": ctx.numa_node, "host_cpus_of_rank0": len(os.sched_getaffinity(0)),
                "public_constructor_wall_ms": public_ms, "public_constructor_first_call_ms": public_first_ms,
                "public_constructor_what": "wall clock of ossg.TripolarGrid(GPU(), size=...) per call, allocation of the 20 arrays and "
                                           "the workspace included, device synchronised; mean of 5 calls after 2 (first_call: the arrays come from cudaMalloc)"},
        "gpu_launches": args.steps * 3,      # tables_kernel, coords(_pair)_kernel, metrics_tile_kernel (the halo frame rides in its first blocks)
        "parity": parity,
        "clocks": clocks,
        "zipper": zipper,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:               # noqa: BLE001  (the line is out; a teardown error changes nothing)
            pass
    if parity["mismatches"]:
        sys.exit(3)


# FLUSHED configuration first (the headline of the zipper roofline), the repeat loop beside it; `traffic` = DRAM bytes per
# launch from the committed ncu captures of the SAME loops (tools/prof_fill.py; profiles/r02_fill_dram.md)
NCU_TRAFFIC_ZIPPER_FLUSH4 = 119344960       # dram__bytes_read + write per launch, mean of 4 captured launches (r02b)
NCU_TRAFFIC_ZIPPER_LOOP8 = 237173888
NCU_TRAFFIC_FILL_FLUSH4 = 320533504


def zipper_roofline(zipper, hbm_peak, peak_src):
    """headline: c,u,v,zeta per launch with the L2 flushed by a 256 MB write before every timed launch (the
    conservative number: the flush's dirty lines are written back during the timed kernel); beside it the same
    after a clean flush, the 8-field repeat loop and the whole fill_halo_regions!"""
    head = zipper.get("batch4")
    if not head or "gbs" not in head:
        return None
    out = {"bound": "hbm", "kernel": "zipper_kernel", "achieved": head["gbs"], "peak": hbm_peak, "unit": "GB/s",
           "frac": head["gbs"] / hbm_peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": head["algorithmic_bytes"],
           "traffic": NCU_TRAFFIC_ZIPPER_FLUSH4, "traffic_source": "profiles/r02_fill_dram.md (ncu dram__bytes of the same loops)",
           "measured_on": "batch4: c,u,v,zeta per launch, L2 flushed (256 MB write) before every timed launch"}
    for key, traffic in (("batch4_clean_flush", NCU_TRAFFIC_ZIPPER_FLUSH4), ("repeat100_zipper_batch8_larger_than_l2", NCU_TRAFFIC_ZIPPER_LOOP8),
                         ("fill_halo_regions_batch4", NCU_TRAFFIC_FILL_FLUSH4)):
        z = zipper.get(key)
        if z and "gbs" in z:
            out[key] = {"achieved": z["gbs"], "frac": z["gbs"] / hbm_peak, "algorithmic_bytes_per_launch": z["algorithmic_bytes"],
                        "traffic": traffic, "dram_gbs": traffic / (z["ms"] * 1e-3) / 1e9, "dram_frac": traffic / (z["ms"] * 1e-3) / 1e9 / hbm_peak}
    return out


def bench_small_configs(ctx):
    """BASELINE.json configs A, B, C (60x30, 360x180x50, 1440x720x75; Float64 and Float32): device time of the full
    construction and of fill_halo_regions! of c,u,v,zeta, next to the CPU restatement of the reference on one thread
    (step for step like the reference's own kernels) and on all host threads."""
    import numpy as np
    from oracle import oracle as orc
    from ossg_b200.grids import make_params
    torch, _lib, L, W = ctx.torch, ctx._lib, ctx.L, WORKLOAD
    out = {}
    o = orc.Oracle(libm=True)
    H = W["halo"]
    for name, (Nx, Ny, Nz) in (("A_60x30x1", (60, 30, 1)), ("B_360x180x50", (360, 180, 50)), ("C_1440x720x75", (1440, 720, 75))):
        res = {}
        p = make_params((Nx, Ny, Nz), (H, H, H), W["south"], W["north_poles_latitude"], W["first_pole_longitude"], W["radius"])
        for et, tdt, tag in ((_lib.OSSG_F64, torch.float64, "f64"), (_lib.OSSG_F32, torch.float32, "f32")):
            shape = (Ny + 2 * H, Nx + 2 * H)
            block = torch.empty((20,) + shape, dtype=tdt, device=ctx.dev)
            optr = _lib.ptr_array([block[k].data_ptr() for k in range(20)])
            nb = L.ossg_tripolar_workspace_bytes(C.byref(p), et)
            ws = torch.empty(max(nb, 128), dtype=torch.uint8, device=ctx.dev)
            _lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), ctx.stream.cuda_stream))

            def step():
                _lib.check(L.ossg_tripolar_build(C.byref(p), et, optr, ws.data_ptr(), ws.numel(), ctx.stream.cuda_stream))

            def cold():
                _lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), ctx.stream.cuda_stream))
                step()
            t = ctx.timed(step, 50, warm=5)
            tc = ctx.timed(cold, 50, warm=5)
            res[f"construct_{tag}"] = {"ms": t, "cells_per_s": Nx * Ny / (t * 1e-3), "first_call_ms": tc}
            if tag == "f64":
                # the same build captured into a CUDA graph and replayed (the three launches of a build are enqueued on the
                # caller's stream only, so it is capturable): what a caller that rebuilds small grids often would do
                try:
                    cs = torch.cuda.Stream(device=ctx.dev)          # capture needs a non-default stream

                    def gstep():
                        _lib.check(L.ossg_tripolar_build(C.byref(p), et, optr, ws.data_ptr(), ws.numel(), cs.cuda_stream))
                    torch.cuda.synchronize(ctx.dev)
                    with torch.cuda.stream(cs):
                        gstep()
                    torch.cuda.synchronize(ctx.dev)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=cs):
                        gstep()
                    tg = ctx.timed(graph.replay, 50, warm=5)
                    res["construct_f64"]["graph_replay_ms"] = tg
                    del graph
                except Exception as e:      # never fatal for the benchmark line
                    res["construct_f64"]["graph_replay_error"] = f"{type(e).__name__}: {e}"[:200]
            # fill_halo_regions! of c,u,v,zeta, repeated back to back (BASELINE config 3: "repeated fill_halo_regions!")
            fshape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
            specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
            fields = [torch.zeros(fshape, dtype=tdt, device=ctx.dev) for _ in specs]
            descs = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(f.data_ptr(), et, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, sg)
                                            for f, (loc, sg) in zip(fields, specs)])
            esz = 8 if et == _lib.OSSG_F64 else 4
            algo = sum(2 * H * Nx * Nz * esz + (Nx * Nz * esz if loc[1] == 0 else 0) for loc, _ in specs)
            algo_full = algo + 4 * (Ny + 2 * H) * (Nz + 2 * H) * 2 * (2 * H) * esz
            tz = ctx.timed(lambda: _lib.check(L.ossg_zipper_fill_batch(descs, 4, ctx.stream.cuda_stream)), 200, warm=5)
            tf = ctx.timed(lambda: _lib.check(L.ossg_fill_halo_regions_batch(descs, 4, ctx.stream.cuda_stream)), 200, warm=5)
            res[f"zipper_batch4_{tag}"] = {"us": tz * 1e3, "algorithmic_bytes": algo, "gbs": algo / (tz * 1e-3) / 1e9}
            res[f"fill_halo_regions_batch4_{tag}"] = {"us": tf * 1e3, "algorithmic_bytes": algo_full, "gbs": algo_full / (tf * 1e-3) / 1e9}
            del fields, block, ws
        # CPU restatement: one thread (a default `julia` session), then all host threads
        po = orc.make_params((Nx, Ny, 1), (H, H, H), W["south"], W["north_poles_latitude"], W["first_pole_longitude"], W["radius"])
        for threads, tag in ((1, "cpu_1_thread"), (os.cpu_count() or 1, "cpu_all_threads")):
            o.set_threads(threads)
            o.build(po)
            reps = 3 if Nx * Ny < 500000 else 1
            t0 = time.perf_counter()
            for _ in range(reps):
                o.build(po)
            dt = (time.perf_counter() - t0) / reps
            res[tag] = {"ms": dt * 1e3, "cells_per_s": Nx * Ny / dt, "threads": threads, "kind": "port"}
        out[name] = res
    return out


class _RawField:
    """adapter: a bare tensor + descriptor where FillReference expects a Field"""

    def __init__(self, data, desc):
        self.data, self._desc = data, desc

    def descriptor(self):
        return self._desc


def bench_zipper(ctx, L, _lib, torch, dev, stream, W, nx, ny):
    """zipper halo-fill GB/s: c (C,C,+1), u (F,C,-1), v (C,F,-1), zeta (F,F,+1) on the workload grid,
    FP64, one batched launch per fill, L2 flushed between timed fills."""
    H, Nz = W["halo"], W["Nz"]
    shape = (Nz + 2 * H, ny + 2 * H, nx + 2 * H)
    specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
    try:
        fields = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in specs]
    except RuntimeError as e:                                   # not enough memory for 4 x 8 GB
        return {"error": str(e)[:100]}
    descs = (_lib.FieldDesc * len(specs))(*[
        _lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, nx, ny, Nz, H, H, H, loc[0], loc[1], 0, 0, s)
        for t, (loc, s) in zip(fields, specs)])
    for k in range(len(specs)):
        _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(descs[k]), 0x5EED + k, 1, 1, stream.cuda_stream))
    # ---- parity at full size first: ONE zipper fill, then (fresh fields) ONE fill_halo_regions! of c,u,v,zeta against the
    # oracle's fill of the global field regenerated on the host (numpy hash), both boundary levels, whole parent window
    res = {}
    try:
        ref = FillReference(ctx, 4)
        raw = [_RawField(t, descs[k]) for k, t in enumerate(fields)]
        par = {"what": "4320x2160 fields c,u,v,zeta, levels 1 and Nz, against the oracle on the host-regenerated global field: the zipper "
                       "over all rows x interior columns, fill_halo_regions! over every element of the padded window (x halos, corners)"}
        for name, fn, full in (("zipper", L.ossg_zipper_fill_batch, False), ("fill_halo_regions", L.ossg_fill_halo_regions_batch, True)):
            for k in range(len(specs)):
                _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(descs[k]), 0x5EED + k, 1, 1, stream.cuda_stream))
            _lib.check(fn(descs, 4, stream.cuda_stream))
            torch.cuda.synchronize(dev)
            par[name] = {"mismatches": ref.mismatches(raw, (1, 1, nx, ny), full=full)}
        ref.close()
        par["mismatches"] = sum(v["mismatches"] for v in par.values() if isinstance(v, dict))
        res["parity"] = par
    except Exception as e:          # noqa: BLE001  (the timing below does not depend on it; a failure is reported, not hidden)
        res["parity"] = {"error": f"{type(e).__name__}: {e}"[:300], "mismatches": 1}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    algo = 0
    for (loc, _) in specs:
        algo += 2 * H * nx * Nz * 8 + (nx * Nz * 8 if loc[1] == 0 else 0)
    # periodic west/east fill over the whole (y, z) parent extent: read + write of 2 Hx elements per row
    rows = (ny + 2 * H) * (Nz + 2 * H)
    algo_full = algo + len(specs) * rows * 2 * (2 * H) * 8
    for name, n, fn in (("batch4", 4, L.ossg_zipper_fill_batch), ("single", 1, L.ossg_zipper_fill_batch),
                        ("fill_halo_regions_batch4", 4, L.ossg_fill_halo_regions_batch)):
        times = []
        for it in range(8):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(fn(descs, n, stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        b = algo_full if name.startswith("fill") else algo if n == 4 else (2 * H * nx * Nz * 8 + nx * Nz * 8)
        res[name] = {"ms": ms, "algorithmic_bytes": b, "gbs": b / (ms * 1e-3) / 1e9}
    res["fill_halo_regions_batch4"]["what"] = ("zipper north fill + periodic west/east fill of c,u,v,zeta in ONE launch "
                                               "(fill_fused_kernel); the periodic part moves 32-byte pieces, one DRAM sector each way")
    # repeated fill_halo_regions! on the same fields, back to back, no flush (SURVEY 8d "repeated fill"):
    # the 117 MB the zipper touches stay resident in the 126 MB L2, the loop is launch/latency bound
    for name, fn, b in (("repeat200_zipper_batch4_no_flush", L.ossg_zipper_fill_batch, algo),
                        ("repeat200_fill_halo_regions_batch4_no_flush", L.ossg_fill_halo_regions_batch, algo_full)):
        for _ in range(5):
            _lib.check(fn(descs, 4, stream.cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(200):
            _lib.check(fn(descs, 4, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 200
        res[name] = {"ms": ms, "algorithmic_bytes": b, "gbs": b / (ms * 1e-3) / 1e9}
    res["config"] = f"{nx}x{ny}x{Nz} FP64 fields c,u,v,zeta, Hy=4, L2 flushed (256 MB write) before each timed fill"
    # ---- the same batched fill with a CLEAN flush: the 256 MB write above leaves the L2 full of dirty
    # lines whose write-back shares the DRAM with the timed kernel; reading a second 256 MB buffer
    # afterwards replaces them by clean lines (the inputs are still not in the L2)
    flush2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush2.zero_()
    times = []
    for it in range(8):
        flush.fill_(it)
        sink = flush2.view(torch.int64).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.ossg_zipper_fill_batch(descs, 4, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    res["batch4_clean_flush"] = {"ms": ms, "algorithmic_bytes": algo, "gbs": algo / (ms * 1e-3) / 1e9,
                                 "what": "L2 flushed by a 256 MB write followed by a 256 MB read (no dirty lines left to write back "
                                         "during the timed fill)"}
    del flush2, sink
    # ---- 8 fields per launch (c,u,v,zeta twice: e.g. two time levels), repeated back to back WITHOUT a
    # flush: one fill touches 235 MB, more than the 126 MB L2, so every fill streams from and to HBM
    # ("inputs larger than L2"); this is the sustained rate a model's halo-fill loop sees
    try:
        more = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in specs]
    except RuntimeError as e:                                   # not enough memory for 8 x 8 GB
        more = None
        res["repeat100_zipper_batch8_larger_than_l2"] = {"error": str(e)[:100]}
    if more is not None:
        descs8 = (_lib.FieldDesc * 8)(*[
            _lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, nx, ny, Nz, H, H, H, loc[0], loc[1], 0, 0, s)
            for t, (loc, s) in zip(fields + more, specs + specs)])
        for k in range(4, 8):
            _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(descs8[k]), 0x5EED + k, 1, 1, stream.cuda_stream))
        for _ in range(5):
            _lib.check(L.ossg_zipper_fill_batch(descs8, 8, stream.cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(100):
            _lib.check(L.ossg_zipper_fill_batch(descs8, 8, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 100
        res["repeat100_zipper_batch8_larger_than_l2"] = {
            "ms": ms, "algorithmic_bytes": 2 * algo, "gbs": 2 * algo / (ms * 1e-3) / 1e9,
            "what": "8 fields per launch, 100 fills back to back, no flush: 235 MB per fill exceed the 126 MB L2"}
        del more
    del fields
    return res


# ------------------------------------------------------------------ distributed fields (N > 1)
class _StubGrid:
    """only what Field / the exchange classes read from a grid (the construction arrays are not needed here)"""


def make_fields(ctx, Rx, Ry, nfields):
    """c, u, v, zeta (repeated) on this rank's tile of the workload grid, synthetic values by global index"""
    import ossg_b200 as ossg
    from ossg_b200.distributed import tile_of
    torch = ctx.torch
    W = WORKLOAD
    Nx, Ny, Nz, H = W["Nx"], W["Ny"], W["Nz"], W["halo"]
    arch = ossg.Distributed(ossg.GPU(ctx.dev.index), ossg.Partition(Rx, Ry))
    i0, j0, nx, ny = tile_of(Nx, Ny, Rx, Ry, arch.rx, arch.ry)
    g = _StubGrid()
    g.Nx, g.Ny, g.Nz, g.Hx, g.Hy, g.Hz = nx, ny, Nz, H, H, H
    g.architecture, g.dtype, g.arrays, g.i0, g.j0, g.global_size = arch, torch.float64, {}, i0, j0, (Nx, Ny, Nz)
    g.size, g.halo = (nx, ny, Nz), (H, H, H)
    C_, F_ = ossg.Center, ossg.Face
    locs = ([(C_, C_, C_), (F_, C_, C_), (C_, F_, C_), (F_, F_, C_)] * ((nfields + 3) // 4))[:nfields]
    shape = (Nz + 2 * H, ny + 2 * H, nx + 2 * H)
    fields = [ossg.Field(loc, g, data=torch.empty(shape, dtype=torch.float64, device=ctx.dev)) for loc in locs]
    return fields, arch, (i0, j0, nx, ny)


def synth(ctx, fields, tile):
    for k, f in enumerate(fields):
        d = f.descriptor()
        ctx._lib.check_util(ctx.U.ossg_synth_field(C.byref(d), 0x5EED + k, tile[0], tile[1], ctx.stream.cuda_stream))


class FillReference:
    """The expected result of ONE fill of the synthetic fields at the first and the last interior level: the GLOBAL
    field regenerated on the host by an independent numpy restatement of the hash (host_synth_levels), filled by the
    oracle (zipper, or zipper + periodic x = fill_halo_regions!), of which every rank owns a window
    (src/distributed_tripolar_grid.jl:112-120).  Rank 0 generates, everybody maps."""

    def __init__(self, ctx, nfields):
        import numpy as np
        W = WORKLOAD
        self.ctx, self.ks = ctx, [1, W["Nz"]]
        d = f"/dev/shm/ossg_fill_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if ctx.world > 1 else os.getpid()}"
        if ctx.rank == 0:
            os.makedirs(d, exist_ok=True)
            for k in range(nfields):
                np.save(os.path.join(d, f"f{k}.npy"), host_synth_levels(0x5EED + k, W["Nx"], W["Ny"], W["halo"], self.ks))
        box = [d]
        if ctx.world > 1:
            ctx.dist.broadcast_object_list(box, src=0)
        self.dir = box[0]
        ctx.barrier()
        self.base = [np.load(os.path.join(self.dir, f"f{k}.npy"), mmap_mode="r") for k in range(nfields)]
        self.cache = {}

    def expected(self, k, f, full):
        import numpy as np
        from oracle import oracle as orc
        key = (k, full)
        if key not in self.cache:
            W = WORKLOAD
            o = FillReference._oracle = getattr(FillReference, "_oracle", None) or orc.Oracle()
            c = np.array(self.base[k])                                   # (2, Ny+2H, Nx+2H): two levels, no z halo
            d = f.descriptor()
            fn = o.fill_halo_regions if full else o.zipper
            fn(c, W["Nx"], W["Ny"], 2, W["halo"], W["halo"], 0, d.locx, d.locy, d.sign)
            self.cache[key] = c
        return self.cache[key]

    def mismatches(self, fields, tile, full):
        """differing bit patterns between this rank's tiles (two levels per field, whole window: interior and
        every halo) and the oracle's global result"""
        import numpy as np
        W = WORKLOAD
        H, Nz = W["halo"], W["Nz"]
        i0, j0, nx, ny = tile
        bad = 0
        # a fold alone leaves the x halos of a tile untouched (they hold neighbours' interior columns in the global
        # field): compare the interior columns of every row then; a whole fill is compared over the whole window
        cols = slice(None) if full else slice(H, H + nx)
        for k, f in enumerate(fields):
            want = self.expected(k, f, full)[:, j0 - 1:j0 - 1 + ny + 2 * H, i0 - 1:i0 - 1 + nx + 2 * H][:, :, cols]
            got = self.ctx.torch.stack([f.data[H + kk - 1] for kk in self.ks]).cpu().numpy()[:, :, cols]
            bad += int((np.ascontiguousarray(got).view(np.uint64) != np.ascontiguousarray(want).view(np.uint64)).sum())
        return bad

    def close(self):
        self.ctx.barrier()
        if self.ctx.rank == 0:
            for n in os.listdir(self.dir):
                os.remove(os.path.join(self.dir, n))
            os.rmdir(self.dir)


# ------------------------------------------------------------------ fold exchange (N > 1)
def bench_fold_exchange(ctx):
    """Fold exchange of c,u,v,zeta (FP64, Nz=100, Hy=4) on the 1/12 degree grid split (N,1): the NCCL arm of the C ABI
    (ossg_fold_exchange: pack kernel, one ncclGroup, unpack kernel) against ONE kernel per rank over NVLink peer memory
    (pull, push, bulk push).  Every form is first checked against the oracle on freshly generated fields."""
    from ossg_b200.distributed import NcclExchange, PeerFold
    world = ctx.world
    H, Nz = WORKLOAD["halo"], WORKLOAD["Nz"]
    all_fields, arch, tile = make_fields(ctx, world, 1, 16)
    nx = tile[2]
    fields = all_fields[:4]
    reps = 20

    def recv_bytes(fs):
        # bytes each rank receives per exchange: Hy rows (+1 for y-Center) x nx columns x Nz levels per field
        # (the row-Ny overwrite only concerns the right half of the globe: half the ranks on average)
        return sum((H + (0.5 if f.descriptor().locy == 0 else 0)) * nx * Nz * 8 for f in fs)

    ref = FillReference(ctx, 4)
    parity = {"what": "one exchange of freshly generated c,u,v,zeta; both boundary levels of every tile, all rows (interior, north halo, "
                      "south halo) x interior columns, against oracle.zipper of the global field regenerated on the host (numpy hash)",
              "partition": f"({world},1)"}
    out = {"partition": f"({world},1)", "fields": 4, "recv_bytes_per_rank": recv_bytes(fields)}

    # ---- NCCL arm (library path)
    nccl = NcclExchange(fields)
    synth(ctx, fields, tile)
    nccl.fold()
    ctx.torch.cuda.synchronize(ctx.dev)
    parity["nccl"] = {"mismatches": ctx.sum_over_ranks(ref.mismatches(fields, tile, full=False))}
    res = {}
    for nf in (4, 8, 16):
        ex = NcclExchange(all_fields[:nf])
        t = ctx.timed(ex.fold, reps)
        rb = recv_bytes(all_fields[:nf])
        res[f"fields{nf}"] = {"ms_per_exchange": t, "recv_bytes_per_rank": rb, "gbs_per_direction_per_gpu": rb / (t * 1e-3) / 1e9}
        del ex
    res["what"] = "ossg_fold_exchange: one pack kernel, one ncclGroup (one ncclSend + ncclRecv per partner), one unpack kernel; persistent buffer"
    out["nccl_pack_send_unpack"] = res

    # ---- the same exchange as ONE kernel per rank over NVLink peer memory; 1 ... 16 fields per
    # launch separate the fixed cost (launch + flag round trips) from the bandwidth
    for push in (False, True, "bulk"):
        res = {}
        key_out = "peer_kernel_push_bulk" if push == "bulk" else "peer_kernel_push" if push else "peer_kernel_pull"
        try:
            pf = PeerFold(fields, push=push)
            synth(ctx, fields, tile)
            ctx.barrier()
            pf()
            ctx.torch.cuda.synchronize(ctx.dev)
            parity[key_out] = {"mismatches": ctx.sum_over_ranks(ref.mismatches(fields, tile, full=False))}
            pf.close()
            for nf in (1, 2, 4, 8, 16):
                pf = PeerFold(all_fields[:nf], push=push)
                t = ctx.timed(pf, reps)
                rb = recv_bytes(all_fields[:nf])
                res[f"fields{nf}"] = {"ms_per_exchange": t, "recv_bytes_per_rank": rb, "gbs_per_direction_per_gpu": rb / (t * 1e-3) / 1e9}
                pf.close()
        except Exception as e:          # noqa: BLE001  (one form failing must not drop the others)
            res["error"] = f"{type(e).__name__}: {e}"[:300]
            out[key_out] = res
            continue
        t1, t8 = res["fields1"], res["fields16"]
        slope = (t8["ms_per_exchange"] - t1["ms_per_exchange"]) / (t8["recv_bytes_per_rank"] - t1["recv_bytes_per_rank"])
        res["marginal_gbs"] = 1e-6 / slope
        res["fixed_us"] = 1e3 * (t1["ms_per_exchange"] - slope * t1["recv_bytes_per_rank"])
        # NVLink 5 nominal 900 GB/s per direction; 770 GB/s is what the driver's own copy probe measured on this pool
        # (VERDICT round 1), 731-741 what a cudaMemcpyPeer reaches (tools/p2p_probe.py, profiles/r01_p2p_probe.txt)
        for key in ("fields4", "fields8", "fields16"):
            res[key]["frac_of_nvlink_900"] = res[key]["gbs_per_direction_per_gpu"] / 900.0
            res[key]["frac_of_measured_770"] = res[key]["gbs_per_direction_per_gpu"] / 770.0
        out[key_out] = res
    # ---- copy-engine form, data phase only: one strided 3-D peer copy per field pulls the
    # mirror rank's source rows (Hy or Hy + 1 rows x nx columns x Nz levels) into a local staging buffer.  Not timed: the
    # local unpack-with-reversal kernel that would follow (HBM speed) and the flag protocol -- so this is an upper bound
    # of what a copy-engine fold could reach, to be held against the kernels' marginal bandwidth above.
    try:
        out["copy_engine_pull_data_phase"] = bench_copy_engine_pull(ctx, all_fields, tile, recv_bytes, reps)
    except Exception as e:              # noqa: BLE001
        out["copy_engine_pull_data_phase"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    ref.close()
    parity["mismatches"] = sum(v["mismatches"] for v in parity.values() if isinstance(v, dict))
    out["parity"] = parity
    return out


def bench_copy_engine_pull(ctx, all_fields, tile, recv_bytes, reps):
    from cuda.bindings import runtime as cudart
    from ossg_b200.distributed import PeerFold
    torch = ctx.torch
    H, Nz = WORKLOAD["halo"], WORKLOAD["Nz"]
    i0, j0, nx, ny = tile
    Ny = WORKLOAD["Ny"]
    res = {}
    for nf in (1, 4, 8, 16):
        fs = all_fields[:nf]
        pf = PeerFold(fs, push=False)                    # (used for its CUDA IPC mappings of the partners' fields only)
        mirror = pf.arch.Rx - 1 - pf.arch.rx
        staging = [torch.empty((Nz, H + 1, nx), dtype=torch.float64, device=ctx.dev) for _ in fs]
        pitch = (nx + 2 * H) * 8
        parms = []
        for k, f in enumerate(fs):
            locy = f.descriptor().locy
            rows = H + (1 if locy == 0 else 0)
            j_first = Ny - H + (1 if locy else 0)            # global row of the first source row (y-Center: Ny-Hy .. Ny, y-Face: Ny-Hy+1 .. Ny)
            row0 = (j_first - j0) + H                        # its index in the partner's parent array (same j0 on the whole top row)
            src = pf.peer_ptrs[k][mirror] + (H * (ny + 2 * H) * (nx + 2 * H) + row0 * (nx + 2 * H) + H) * 8
            pm = cudart.cudaMemcpy3DParms()
            pm.srcPtr = cudart.make_cudaPitchedPtr(src, pitch, nx + 2 * H, ny + 2 * H)
            pm.dstPtr = cudart.make_cudaPitchedPtr(staging[k].data_ptr(), nx * 8, nx, H + 1)
            pm.srcPos = cudart.make_cudaPos(0, 0, 0)
            pm.dstPos = cudart.make_cudaPos(0, 0, 0)
            pm.extent = cudart.make_cudaExtent(nx * 8, rows, Nz)
            pm.kind = cudart.cudaMemcpyKind.cudaMemcpyDefault
            parms.append(pm)
        sh = ctx.stream.cuda_stream

        def exchange():
            for pm in parms:
                err, = cudart.cudaMemcpy3DAsync(pm, sh)
                if err != cudart.cudaError_t.cudaSuccess:
                    raise RuntimeError(f"cudaMemcpy3DAsync: {err}")
        t = ctx.timed(exchange, reps)
        rb = recv_bytes(fs)
        res[f"fields{nf}"] = {"ms_per_exchange": t, "recv_bytes_per_rank": rb, "gbs_per_direction_per_gpu": rb / (t * 1e-3) / 1e9}
        # the copies really moved the partner's rows: level 0, first source row against the partner's own value via NCCL-free check
        torch.cuda.synchronize(ctx.dev)
        del staging
        pf.close()
    t1, t16 = res["fields1"], res["fields16"]
    slope = (t16["ms_per_exchange"] - t1["ms_per_exchange"]) / (t16["recv_bytes_per_rank"] - t1["recv_bytes_per_rank"])
    res["marginal_gbs"] = 1e-6 / slope
    res["fixed_us"] = 1e3 * (t1["ms_per_exchange"] - slope * t1["recv_bytes_per_rank"])
    res["what"] = ("one cudaMemcpy3DAsync per field (rows x nx columns x Nz levels, row pitch of the partner's array) from the mirror rank's "
                   "IPC-mapped field into a local staging buffer; no unpack kernel, no flags: data phase of a copy-engine fold only")
    return res


def bench_distributed_fill(ctx):
    """fill_halo_regions! of c,u,v,zeta (FP64, Nz=100, halo 4) on the 1/12 degree grid in an (x, y)
    partition -- the distributed_bickley_jet-style halo-fill loop (BASELINE config 5): fold on the
    top row + south/north rows + west/east columns.  NCCL arm of the C ABI (ossg_fold_exchange +
    ossg_halo_exchange) against the NVLink peer-memory form (three or four kernels per fill, PeerHalo);
    the eager torch.distributed form of round 1 stays as a third line."""
    from ossg_b200.distributed import NcclExchange, PeerHalo, distributed_fill_halo_regions
    world = ctx.world
    H, Nz = WORKLOAD["halo"], WORKLOAD["Nz"]
    Rx, Ry = mixed_partition(world)
    fields, arch, tile = make_fields(ctx, Rx, Ry, 4)
    i0, j0, nx, ny = tile
    reps = 10
    ref = FillReference(ctx, 4)
    parity = {"what": "one fill of freshly generated c,u,v,zeta; both boundary levels of every tile, whole window (interior, fold, "
                      "south/north and west/east halos, corners), against oracle.fill_halo_regions of the global field", "partition": f"({Rx},{Ry})"}
    nccl = NcclExchange(fields)
    peer = PeerHalo(fields)
    for name, fn in (("nccl", nccl), ("peer_memory", peer), ("torch_eager_nccl", lambda: distributed_fill_halo_regions(fields))):
        synth(ctx, fields, tile)
        ctx.barrier()
        fn()
        ctx.torch.cuda.synchronize(ctx.dev)
        parity[name] = {"mismatches": ctx.sum_over_ranks(ref.mismatches(fields, tile, full=True))}
    ref.close()
    t_nccl = ctx.timed(nccl, reps)
    t_peer = ctx.timed(peer, reps)
    t_eager = ctx.timed(lambda: distributed_fill_halo_regions(fields), 5)
    peer.close()
    # bytes a rank receives per fill: fold rows (top row only), 2 x Hy rows, 2 x Hx columns over the parent extent
    levels, sx, sy = Nz + 2 * H, nx + 2 * H, ny + 2 * H
    recv = len(fields) * 8 * ((2 if Ry > 1 else 0) * H * sx * levels * (0.5 if Ry == 2 else 1.0) + 2 * H * sy * levels)
    if arch.is_top_row():
        recv += sum((H + (0.5 if f.descriptor().locy == 0 else 0)) * nx * Nz * 8 for f in fields)
    parity["mismatches"] = sum(v["mismatches"] for v in parity.values() if isinstance(v, dict))
    return {"partition": f"({Rx},{Ry})", "fields": len(fields), "approx_recv_bytes_top_row_rank": recv,
            "nccl": {"ms_per_fill": t_nccl, "what": "ossg_fold_exchange + ossg_halo_exchange (C ABI): per phase one pack kernel, one ncclGroup, one unpack kernel"},
            "peer_memory": {"ms_per_fill": t_peer, "kernels_per_fill": 4 if arch.is_top_row() else 3},
            "torch_eager_nccl": {"ms_per_fill": t_eager, "what": "round-1 arm: eager slicing + batch_isend_irecv + host waits"},
            "speedup_peer_over_nccl": t_nccl / t_peer, "parity": parity}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
