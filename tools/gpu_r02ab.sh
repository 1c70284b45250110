#!/bin/bash
# round-2 GPU call AB (1 GPU): table kernel launched with programmatic serialisation behind the previous build's tile kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do
for v in head cur; do
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/variant_time.py 2>&1 | tail -1
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
done
done | tee gpurun_out/r02ab_pdl0.log
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02ab_pytest_poison.log 2>&1
tail -6 gpurun_out/r02ab_pytest_poison.log
