#!/bin/bash
# round-2 GPU call O (1 GPU): tests, variants (integer octant tests in the point-stage trig vs the last commit; coordinate loop split by x location)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02o_pytest.log 2>&1
tail -4 gpurun_out/r02o_pytest.log
for rep in 1 2; do
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02o_variants.log
for v in cur cxsplit; do
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py --xsplit 2>&1 | tail -1
done | tee gpurun_out/r02o_slabs.log
