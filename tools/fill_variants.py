"""Times the forms of fill_halo_regions! (zipper + periodic west/east fill of c,u,v,zeta, FP64, 1/12 degree,
Nz = 100) selected by OSSG_FILL_VARIANT (csrc/ossg_zipper.cu, fill_variant()) with the L2 flushed before
every timed launch, and prints one line per variant with a checksum of the filled fields (all variants
must agree bit for bit).  Optionally sets the device's L2 fetch granularity first.

    python tools/fill_variants.py                       # parent: runs every variant in its own process
    python tools/fill_variants.py --one V [--gran G]    # child
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [0, 1, 3, 31, 33, 11]      # (+30: L2::64B prefetch-size hint on the piece loads; form 0 has no 32-byte-piece loads to hint)


def child(variants, gran):
    import torch
    import ossg_b200  # noqa: F401
    from ossg_b200 import _lib
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    got = None
    if gran:
        rt = C.CDLL("libcudart.so.12")
        rc = rt.cudaDeviceSetLimit(5, C.c_size_t(gran))          # cudaLimitMaxL2FetchGranularity
        v = C.c_size_t()
        rt.cudaDeviceGetLimit(C.byref(v), 5)
        got = (rc, v.value)
    L = _lib.lib()
    Nx, Ny, Nz, H = 4320, 2160, 100, 4
    shape = (Nz + 2 * H, Ny + 2 * H, Nx + 2 * H)
    specs = [((0, 0), 1.0), ((1, 0), -1.0), ((0, 1), -1.0), ((1, 1), 1.0)]
    fields = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in specs]
    descs = (_lib.FieldDesc * 4)(*[_lib.FieldDesc(t.data_ptr(), _lib.OSSG_F64, Nx, Ny, Nz, H, H, H, loc[0], loc[1], 0, 0, s)
                                    for t, (loc, s) in zip(fields, specs)])
    s = torch.cuda.current_stream(dev)
    for k in range(4):
        _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(descs[k]), 0x5EED + k, 1, 1, s.cuda_stream))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for variant in variants:
        L.ossg_experiment_fill_variant(int(variant))
        for k in range(4):
            _lib.check_util(_lib.testutil().ossg_synth_field(C.byref(descs[k]), 0x5EED + k, 1, 1, s.cuda_stream))
        one(variant, got, torch, _lib, L, dev, s, descs, fields, specs, flush, Nx, Ny, Nz, H)


def one(variant, got, torch, _lib, L, dev, s, descs, fields, specs, flush, Nx, Ny, Nz, H):
    res = {}
    for name, fn in (("fill", L.ossg_fill_halo_regions_batch), ("zipper", L.ossg_zipper_fill_batch)):
        times = []
        for it in range(10):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            _lib.check(fn(descs, 4, s.cuda_stream))
            e1.record(s)
            torch.cuda.synchronize(dev)
            if it >= 3:
                times.append(e0.elapsed_time(e1) * 1e3)
        times.sort()
        res[name + "_us_median"] = times[len(times) // 2]
        res[name + "_us_min"] = times[0]
    chk = 0
    for t in fields:
        chk ^= int(t.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
    algo = sum(2 * H * Nx * Nz * 8 + (Nx * Nz * 8 if loc[1] == 0 else 0) for loc, _ in specs)
    algo_full = algo + 4 * (Ny + 2 * H) * (Nz + 2 * H) * 2 * (2 * H) * 8
    res.update({"variant": variant, "gran": got, "checksum": f"{chk:016x}",
                "fill_gbs": algo_full / res["fill_us_median"] / 1e3, "zipper_gbs": algo / res["zipper_us_median"] / 1e3})
    print(json.dumps(res), flush=True)


def main():
    if "--one" in sys.argv:
        v = [int(x) for x in sys.argv[sys.argv.index("--one") + 1].split(",")]
        g = int(sys.argv[sys.argv.index("--gran") + 1]) if "--gran" in sys.argv else 0
        child(v, g)
        return
    for gran in (0, 32):               # the fetch granularity is a device limit: one process per setting
        vs = VARIANTS if gran == 0 else [1, 31]
        cmd = [sys.executable, __file__, "--one", ",".join(map(str, vs))] + (["--gran", str(gran)] if gran else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        print(r.stdout.strip(), flush=True)
        if r.returncode:
            print(json.dumps({"gran": gran, "error": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
