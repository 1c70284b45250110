#!/bin/bash
# round-2 GPU call G (1 GPU): tests in every frame mode, slab timings (per-rank windows of 1/2/4/8-rank Y partitions) per mode
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in fork tail ride; do
( time OSSG_FRAME=$m timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parity or tile or construct or distributed or window or matches or metrics" 2>&1 | tail -4 ) > gpurun_out/r02g_pytest_$m.log 2>&1
tail -3 gpurun_out/r02g_pytest_$m.log
done
for rep in 1 2; do
OSSG_B200_LIB=$PWD/build/variants/libossg_head.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
for m in fork tail ride; do OSSG_FRAME=$m timeout 120 python tools/slab_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02g_slabs.log
