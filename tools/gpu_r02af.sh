#!/bin/bash
# round-2 GPU call AF (1 GPU): north-halo coordinate rows through an out-of-line straight-line point: timings, then tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do timeout 120 python tools/variant_time.py 2>&1 | tail -1; timeout 120 python tools/slab_time.py 2>&1 | tail -1; done | tee gpurun_out/r02af_times.log
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02af_pytest_poison.log 2>&1
tail -4 gpurun_out/r02af_pytest_poison.log
