"""NVLink store-path probe (run under torchrun, one rank per GPU): every rank copies the same volume
into the memory of its mirror rank (world-1-rank, the fold's partner) with ossg_peer_copy_probe in
several forms -- SM-issued 8/16-byte stores, shared-memory staging + bulk stores, bulk loads + bulk
stores -- onto a contiguous destination and onto fold-like row segments.  Prints GB/s per direction
per GPU (max time over ranks) next to a cudaMemcpyPeer-style torch copy of the same volume."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ossg_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
SPAN = 4 << 30
dst = torch.empty(SPAN + 4096, dtype=torch.uint8, device=dev)
src = torch.empty(160 << 20, dtype=torch.uint8, device=dev).random_(0, 255)
h, off = _lib.IpcHandle(), C.c_uint64()
_lib.check(L.ossg_ipc_export(dst.data_ptr(), C.byref(h), C.byref(off)))
got = [None] * world
dist.all_gather_object(got, (bytes(h.bytes), int(off.value)))
partner = world - 1 - rank
if partner == rank:
    peer = dst.data_ptr()
else:
    hh, out = _lib.IpcHandle(), C.c_void_p()
    C.memmove(hh.bytes, got[partner][0], 64)
    _lib.check(L.ossg_ipc_open(C.byref(hh), got[partner][1], C.byref(out)))
    peer = out.value
stream = torch.cuda.current_stream(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


MB = 1 << 20
layouts = {
    # name: (rows, seg_bytes, group, group_stride, pitch, base offset)
    "contiguous 117 MB":                 (7168, 16384, 1, 16384, 16384, 0),
    "fold rows, 2-GPU tile (17280 B)":   (8000, 17280, 5, 2359296, 17344, 32),
    "fold rows, 8-GPU tile (4320 B)":    (8000 * 4, 4320, 5, 589824, 4384, 32),
}
results = []
for lname, (rows, seg, group, gs, pitch, base) in layouts.items():
    vol = rows * seg
    assert vol <= src.numel() and (rows // group + 1) * gs <= SPAN
    for mode, chunk in ((0, 2048), (1, 4096), (2, 2048), (2, 8192), (2, 16384), (3, 2048), (3, 8192), (3, 16384)):
        for blocks in ((148 * 4,) if mode < 3 else (148 * 4, 148 * 8)):
            def fn():
                _lib.check_util(_lib.testutil().ossg_peer_copy_probe(peer + base, src.data_ptr(), rows, seg, group, gs, pitch, chunk, mode, blocks, stream.cuda_stream))
            ms = timed(fn)
            results.append({"layout": lname, "mode": mode, "chunk": chunk, "blocks": blocks, "bytes": vol, "ms": ms, "gbs": vol / ms / 1e6})
            if rank == 0:
                print(f"{lname:34s} mode {mode} chunk {chunk:5d} blocks {blocks:4d}: {ms * 1e3:8.1f} us  {vol / ms / 1e6:7.1f} GB/s", flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/peer_copy_probe.json", "w"), indent=0)
dist.destroy_process_group()
