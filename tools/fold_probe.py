import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = bench.bench_fold_exchange(world, rank, dev)
if rank == 0:
    for k in ("peer_kernel_pull", "peer_kernel_push"):
        r = out[k]; print(k, {n: round(r[n]["ms_per_exchange"]*1e3,1) for n in ("fields1","fields4","fields8")}, "marginal", round(r["marginal_gbs"]), "fixed", round(r["fixed_us"],1))
dist.destroy_process_group()
