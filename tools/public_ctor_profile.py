"""Where the wall clock of the public constructor goes (never a bench value): cProfile of
ossg.TripolarGrid(GPU(), size=(4320, 2160, 100)) after two warm calls, plus plain timings of its parts."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ossg_b200 as ossg

dev = torch.device("cuda", 0)
size = (4320, 2160, 100)
def call():
    g = ossg.TripolarGrid(ossg.GPU(0), size=size, halo=(4, 4, 4))
    torch.cuda.synchronize(dev)
    return g
for _ in range(2):
    call()
t0 = time.perf_counter()
for _ in range(5):
    call()
print("public constructor, per call: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
g = call()
t0 = time.perf_counter()
for _ in range(5):
    h = call()
print("the same while an earlier grid is alive: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
del g, h
t0 = time.perf_counter()
for _ in range(5):
    out = [torch.empty((2168, 4328), dtype=torch.float64, device=dev) for _ in range(20)]
    del out
torch.cuda.synchronize(dev)
print("20 x torch.empty alone: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
