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
# committed capture profiles/r01i_*_opcodes.md: DFMA + DMUL + DADD + DSETP thread-instructions):
#   metrics_tile_kernel : 915.7 per tile cell (4320 x 2158 cells)
#   coords_pair_kernel  : 194.9 per padded element (4328 x 2168 elements)
# and the DRAM traffic of one metrics_tile_kernel launch from the same capture
# (dram__bytes_read.sum + dram__bytes_write.sum = 601.5 MB + 849.7 MB; algorithmic 1491.6 MB).
# These are properties of the code, not of the timing; refresh them whenever the kernels change.
FP64_INST_PER_CELL_TILE = 915.7
FP64_INST_PER_ELEM_COORDS = 194.9
NCU_TRAFFIC_TILE_BYTES = 601518592 + 849665280
PROFILE_TAG = "r01i"


def partition_for(n):
    """Construction partition at N GPUs: Y slabs, the reference's own (and only) distributed layout
    (src/distributed_tripolar_grid.jl:28-31).  Every rank keeps the full x extent, so the
    coordinate kernel's mirror-pair form applies.  (x, y) partitions are exercised by the parity
    tests (every factorisation of N) and by the fold-exchange benchmark below, which uses (N, 1)."""
    return (1, n)


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
def run_own(args):
    import torch
    import torch.distributed as dist
    import ossg_b200 as ossg
    from ossg_b200 import _lib
    from ossg_b200.distributed import tile_of
    from ossg_b200.grids import make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # (NCCL_DEBUG is left as the environment has it: unset, NCCL prints nothing and stdout is the one JSON line)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; stdout must
        # carry the one JSON line only, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.lib()
    if L.ossg_device_ok() != 0:
        raise SystemExit("no usable CUDA device: " + _lib.last_error())

    W = WORKLOAD
    Rx, Ry = partition_for(world)
    rx, ry = rank // Ry, rank % Ry
    i0, j0, nx, ny = tile_of(W["Nx"], W["Ny"], Rx, Ry, rx, ry)
    H = W["halo"]
    params = make_params((W["Nx"], W["Ny"], W["Nz"]), (H, H, H), W["south"], W["north_poles_latitude"],
                         W["first_pole_longitude"], W["radius"], tile=(i0, j0, nx, ny))
    shape = (ny + 2 * H, nx + 2 * H)
    out = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(20)]
    optr = _lib.ptr_array([t.data_ptr() for t in out])
    nbytes = L.ossg_tripolar_workspace_bytes(C.byref(params), _lib.OSSG_F64)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        _lib.check(L.ossg_tripolar_build(C.byref(params), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_local = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    else:
        ms_total = ms_local
    ms_step = ms_total / args.steps
    cells_global = W["Nx"] * W["Ny"]
    value = cells_global / (ms_step * 1e-3)

    # ---- per-kernel device times (events between the kernels), same workload
    prof = (C.c_float * 4)()
    acc = [0.0] * 4
    nprof = max(3, min(args.steps, 10))
    for _ in range(nprof):
        _lib.check(L.ossg_tripolar_build_profiled(C.byref(params), _lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(),
                                                  stream.cuda_stream, prof))
        for k in range(4):
            acc[k] += prof[k] / nprof
    kernel_ms = dict(zip(["tables", "coords", "metrics_tile", "metrics_frame"], acc))

    # ---- end to end through the public API with host buffers: construct + D2H of the 20 arrays
    host = [torch.empty(shape, dtype=torch.float64).pin_memory() for _ in range(20)]
    pin_params = bytes(params)       # the only input: 64 bytes of scalars, passed as kernel arguments

    def e2e_step():
        step()
        for h, d in zip(host, out):
            h.copy_(d, non_blocking=True)

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    d2h_bytes = 20 * shape[0] * shape[1] * 8
    clocks = sampler.stop() if rank == 0 else None

    # ---- FP64 peak probe (roofline denominator for the FP64-bound kernels)
    ms = C.c_float(); fl = C.c_double()
    _lib.check_util(_lib.testutil().ossg_fp64_peak_probe(20000, C.byref(ms), C.byref(fl), stream.cuda_stream))
    fp64_peak_tflops = fl.value / (ms.value * 1e-3) / 1e12

    # ---- zipper halo fill (second half of the metric): 4 fields c,u,v,zeta on the local tile
    # (last GPU section: whatever happens in the peer-memory plumbing, the construction numbers above stand)
    del out, ws
    torch.cuda.empty_cache()
    if world == 1:
        zipper = bench_zipper(L, _lib, torch, dev, stream, W, nx, ny)
    else:
        # (a failure of the peer-memory plumbing -- e.g. CUDA IPC unavailable in the container -- must not
        # take the construction number down with it)
        zipper = {}
        for key, fn in (("distributed_fill_halo_regions", bench_distributed_fill), ("fold_exchange", bench_fold_exchange)):
            try:
                zipper[key] = fn(world, rank, dev)
            except Exception as e:          # noqa: BLE001
                zipper[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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
            "whole_step_frac": (inst_tile + inst_coords) / (ms_step * 1e-3) / peak_inst,
            "fp64_inst_source": f"ncu SASS opcode counts, profiles/{PROFILE_TAG}_*_opcodes.md"})

    roofline_zipper = None
    if world == 1 and isinstance(zipper, dict):
        best = zipper.get("repeat100_zipper_batch8_larger_than_l2")
        if not best or "gbs" not in best:
            best = zipper.get("batch4")
        if best and "gbs" in best:
            roofline_zipper = {"bound": "hbm", "kernel": "zipper_kernel", "achieved": best["gbs"], "peak": hbm_peak,
                               "unit": "GB/s", "frac": best["gbs"] / hbm_peak, "peak_source": peak_src,
                               "algorithmic_bytes_per_launch": best["algorithmic_bytes"],
                               "measured_on": "repeat100_zipper_batch8_larger_than_l2" if "what" in best and "8 fields" in best["what"] else "batch4"}
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
        "roofline": roofline, "roofline_fp64": roofline_fp64, "roofline_zipper": roofline_zipper,
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                         "sample": cpu_what + ", 3 steps, CPU restatement of reference v0.2.1 (glibc libm); "
                                   "Julia is not installed"},
        "e2e": {"value": cells_global / e2e_s, "unit": UNIT, "h2d_bytes_per_step": len(pin_params),
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s * 1e3,
                "what": "ossg_tripolar_build + copy of the 20 arrays to pinned host buffers (TripolarGrid(CPU()) equivalent)"},
        "gpu_launches": args.steps * 4,
        "clocks": clocks,
        "zipper": zipper,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:               # noqa: BLE001  (the line is out; a teardown error changes nothing)
            pass


def bench_zipper(L, _lib, torch, dev, stream, W, nx, ny):
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    algo = 0
    for (loc, _) in specs:
        algo += 2 * H * nx * Nz * 8 + (nx * Nz * 8 if loc[1] == 0 else 0)
    res = {}
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


# ------------------------------------------------------------------ fold exchange (N > 1)
def bench_fold_exchange(world, rank, dev):
    """Fold exchange of c,u,v,zeta (FP64, Nz=100, Hy=4) on the 1/12 degree grid split (N,1)."""
    import torch
    import torch.distributed as dist
    import ossg_b200 as ossg
    from ossg_b200 import _lib
    from ossg_b200.distributed import PeerFold, fold_exchange, tile_of
    Nx, Ny, Nz, H = 4320, 2160, 100, 4
    arch = ossg.Distributed(ossg.GPU(dev.index), ossg.Partition(world, 1))
    i0, j0, nx, ny = tile_of(Nx, Ny, world, 1, arch.rx, arch.ry)

    class G:        # grid stub: only sizes are needed by Field / fold_exchange
        pass
    g = G()
    g.Nx, g.Ny, g.Nz, g.Hx, g.Hy, g.Hz = nx, ny, Nz, H, H, H
    g.architecture, g.dtype, g.arrays, g.i0, g.j0, g.global_size = arch, torch.float64, {}, i0, j0, (Nx, Ny, Nz)
    g.size, g.halo = (nx, ny, Nz), (H, H, H)
    C_, F_ = ossg.Center, ossg.Face
    locs = [(C_, C_, C_), (F_, C_, C_), (C_, F_, C_), (F_, F_, C_)] * 4      # c, u, v, zeta four times: batches of up to 16
    shape = (Nz + 2 * H, ny + 2 * H, nx + 2 * H)
    all_fields = [ossg.Field(loc, g, data=torch.empty(shape, dtype=torch.float64, device=dev)) for loc in locs]
    fields = all_fields[:4]
    L = _lib.lib()
    for k, f in enumerate(all_fields):
        d = f.descriptor()
        _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(d), 0x5EED + k, i0, j0, torch.cuda.current_stream(dev).cuda_stream))
    for _ in range(3):
        fold_exchange(fields)
    torch.cuda.synchronize(dev)
    dist.barrier()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fold_exchange(fields)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # bytes each rank receives per exchange: Hy rows (+1 for y-Center) x nx columns x Nz levels per field
    # (the row-Ny overwrite only concerns the right half of the globe: half the ranks on average)
    recv = sum((H + (0.5 if f.descriptor().locy == 0 else 0)) * nx * Nz * 8 for f in fields)
    out = {"partition": f"({world},1)", "fields": 4, "recv_bytes_per_rank": recv,
           "nccl_pack_send_unpack": {"ms_per_exchange": float(ms.item()),
                                     "gbs_per_direction_per_gpu": recv / (float(ms.item()) * 1e-3) / 1e9}}
    # ---- the same exchange as ONE kernel per rank over NVLink peer memory; 1, 2 and 4 fields per
    # launch separate the fixed cost (launch + two flag barriers) from the bandwidth
    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a0.elapsed_time(a1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for push in (False, True, "bulk"):
        res = {}
        key_out = "peer_kernel_push_bulk" if push == "bulk" else "peer_kernel_push" if push else "peer_kernel_pull"
        try:
            for nf in (1, 2, 4, 8, 16):
                t = timed(PeerFold(all_fields[:nf], push=push))
                rb = sum((H + (0.5 if f.descriptor().locy == 0 else 0)) * nx * Nz * 8 for f in all_fields[:nf])
                res[f"fields{nf}"] = {"ms_per_exchange": t, "recv_bytes_per_rank": rb, "gbs_per_direction_per_gpu": rb / (t * 1e-3) / 1e9}
        except Exception as e:          # noqa: BLE001  (one form failing must not drop the others)
            res["error"] = f"{type(e).__name__}: {e}"[:300]
            out[key_out] = res
            continue
        t1, t8 = res["fields1"], res["fields16"]
        slope = (t8["ms_per_exchange"] - t1["ms_per_exchange"]) / (t8["recv_bytes_per_rank"] - t1["recv_bytes_per_rank"])
        res["marginal_gbs"] = 1e-6 / slope
        res["fixed_us"] = 1e3 * (t1["ms_per_exchange"] - slope * t1["recv_bytes_per_rank"])
        # NVLink 5 nominal 900 GB/s per direction; 735 GB/s is what a bidirectional cudaMemcpyPeer
        # reaches on this pool (tools/p2p_probe.py, profiles/r01_p2p_probe.txt)
        for key in ("fields4", "fields8", "fields16"):
            res[key]["frac_of_nvlink_900"] = res[key]["gbs_per_direction_per_gpu"] / 900.0
            res[key]["frac_of_measured_copy_735"] = res[key]["gbs_per_direction_per_gpu"] / 735.0
        out[key_out] = res
    return out


def bench_distributed_fill(world, rank, dev):
    """fill_halo_regions! of c,u,v,zeta (FP64, Nz=100, halo 4) on the 1/12 degree grid in an (x, y)
    partition -- the distributed_bickley_jet-style halo-fill loop (BASELINE config 5): fold on the
    top row + south/north rows + west/east columns.  NCCL form (pack/send/recv/unpack through
    torch.distributed) against the NVLink peer-memory form (three kernels per fill, PeerHalo)."""
    import torch
    import torch.distributed as dist
    import ossg_b200 as ossg
    from ossg_b200 import _lib
    from ossg_b200.distributed import PeerHalo, distributed_fill_halo_regions, tile_of
    Nx, Ny, Nz, H = 4320, 2160, 100, 4
    Rx, Ry = {2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
    arch = ossg.Distributed(ossg.GPU(dev.index), ossg.Partition(Rx, Ry))
    i0, j0, nx, ny = tile_of(Nx, Ny, Rx, Ry, arch.rx, arch.ry)

    class G:
        pass
    g = G()
    g.Nx, g.Ny, g.Nz, g.Hx, g.Hy, g.Hz = nx, ny, Nz, H, H, H
    g.architecture, g.dtype, g.arrays, g.i0, g.j0, g.global_size = arch, torch.float64, {}, i0, j0, (Nx, Ny, Nz)
    g.size, g.halo = (nx, ny, Nz), (H, H, H)
    C_, F_ = ossg.Center, ossg.Face
    locs = [(C_, C_, C_), (F_, C_, C_), (C_, F_, C_), (F_, F_, C_)]
    shape = (Nz + 2 * H, ny + 2 * H, nx + 2 * H)
    fields = [ossg.Field(loc, g, data=torch.zeros(shape, dtype=torch.float64, device=dev)) for loc in locs]
    L = _lib.lib()
    for k, f in enumerate(fields):
        d = f.descriptor()
        _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(d), 0x5EED + k, i0, j0, torch.cuda.current_stream(dev).cuda_stream))
    reps = 10

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a0.elapsed_time(a1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_nccl = timed(lambda: distributed_fill_halo_regions(fields))
    t_peer = timed(PeerHalo(fields))
    # bytes a rank receives per fill: fold rows (top row only), 2 x Hy rows, 2 x Hx columns over the parent extent
    levels, sx, sy = Nz + 2 * H, nx + 2 * H, ny + 2 * H
    recv = len(fields) * 8 * ((2 if Ry > 1 else 0) * H * sx * levels * (0.5 if Ry == 2 else 1.0) + 2 * H * sy * levels)
    if arch.is_top_row():
        recv += sum((H + (0.5 if f.descriptor().locy == 0 else 0)) * nx * Nz * 8 for f in fields)
    return {"partition": f"({Rx},{Ry})", "fields": len(fields), "approx_recv_bytes_top_row_rank": recv,
            "nccl": {"ms_per_fill": t_nccl}, "peer_memory": {"ms_per_fill": t_peer, "kernels_per_fill": 3 if arch.is_top_row() else 2},
            "speedup": t_nccl / t_peer}


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
