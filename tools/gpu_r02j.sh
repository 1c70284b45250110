#!/bin/bash
# round-2 GPU call J (1 GPU): tests, tile timings of x-split partitions (pair form of the coordinate kernel for every window) vs the last commit
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02j_pytest.log 2>&1
tail -4 gpurun_out/r02j_pytest.log
for rep in 1 2; do
OSSG_B200_LIB=$PWD/build/variants/libossg_head.so timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1
timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1
timeout 120 python tools/slab_time.py 2>&1 | tail -1
done | tee gpurun_out/r02j_tiles.log
