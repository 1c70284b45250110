#!/bin/bash
# round-2 GPU call K (1 GPU): tests, slab / tile timings with 256- and 128-thread coordinate blocks
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02k_pytest.log 2>&1
tail -4 gpurun_out/r02k_pytest.log
for rep in 1 2; do
for v in cur cb128; do
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/variant_time.py 2>&1 | tail -1
done
done | tee gpurun_out/r02k_tiles.log
