#!/bin/bash
# round-2 GPU call AH (1 GPU): zipper kernel with streaming (evict-first) accesses vs ordinary ones
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2 3; do for v in plain stream; do OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/fill_time.py 2>&1 | tail -1; done; done | tee gpurun_out/r02ah_zipper_stream.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fill or zipper" 2>&1 | tail -2
