#!/bin/bash
# multi-GPU call: N = $1 GPUs, output tag $2.  Distributed parity script, then bench.py --gpus N (with its own parity keys).
N=${1:-2}; TAG=${2:-r02m}
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/run_distributed_gpu.py > gpurun_out/${TAG}_dist_$N.json 2> gpurun_out/${TAG}_dist_$N.err
echo "dist rc=$?" >> gpurun_out/${TAG}_dist_$N.err
tail -c 1500 gpurun_out/${TAG}_dist_$N.json; tail -c 1500 gpurun_out/${TAG}_dist_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_$N.err
tail -c 1500 gpurun_out/${TAG}_bench_$N.err
head -c 600 gpurun_out/${TAG}_bench_$N.json
