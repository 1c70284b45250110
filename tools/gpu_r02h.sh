#!/bin/bash
# round-2 GPU call H (1 GPU): tests, slab timings after the block reorderings
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r02h_pytest.log 2>&1
tail -4 gpurun_out/r02h_pytest.log
for rep in 1 2; do
for m in fork ride; do OSSG_FRAME=$m timeout 120 python tools/slab_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02h_slabs.log
