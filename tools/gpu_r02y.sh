#!/bin/bash
# round-2 GPU call Y (1 GPU): halo images stored by the tile kernel.  GPU suite with poisoned allocations, then the ordinary suite,
# slab / tile / whole-grid timings against the last commit's library
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02y_pytest_poison.log 2>&1
tail -6 gpurun_out/r02y_pytest_poison.log
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02y_pytest.log 2>&1
tail -4 gpurun_out/r02y_pytest.log
for rep in 1 2; do
for v in head cur; do
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/variant_time.py 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1
done
done | tee gpurun_out/r02y_images.log
