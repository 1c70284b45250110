"""NVLink ceiling probe (one process, two GPUs): cudaMemcpyPeer-style copies through torch,
unidirectional and bidirectional, 256 MB per copy."""
import torch
n = 256 << 20
a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
a1 = torch.empty(n, dtype=torch.uint8, device="cuda:1"); b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0 = torch.cuda.Stream(device="cuda:0"); s1 = torch.cuda.Stream(device="cuda:1")

def run(bidir, reps=10):
    for _ in range(2):
        with torch.cuda.stream(s0): a0.copy_(a1, non_blocking=True)       # GPU0 pulls from GPU1
        if bidir:
            with torch.cuda.stream(s1): b1.copy_(b0, non_blocking=True)   # GPU1 pulls from GPU0
    torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s0):
        e0.record(s0)
        for _ in range(reps): a0.copy_(a1, non_blocking=True)
        e1.record(s0)
    if bidir:
        with torch.cuda.stream(s1):
            for _ in range(reps): b1.copy_(b0, non_blocking=True)
    torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("unidirectional GB/s:", run(False))
print("bidirectional GB/s per direction:", run(True))
