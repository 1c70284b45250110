#!/bin/bash
# round-2 GPU call V (1 GPU): fused fill with interleaved part A / part B blocks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do for g in 0 1; do OSSG_FILL_INTERLEAVE=$g timeout 120 python tools/fill_time.py 2>&1 | tail -1 | sed "s/^/interleave=$g /"; done; done | tee gpurun_out/r02v_fill.log
OSSG_FILL_INTERLEAVE=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fill or zipper" 2>&1 | tail -2
OSSG_FILL_INTERLEAVE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fill or zipper" 2>&1 | tail -2
