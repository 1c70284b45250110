#!/bin/bash
# round-2 GPU call L (1 GPU): tests, variants (paired patch points, coordinate block sizes, tile shapes)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02l_pytest.log 2>&1
tail -4 gpurun_out/r02l_pytest.log
OSSG_TILE_PAIRS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair or tile" 2>&1 | tail -3
for rep in 1 2; do
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02l_variants.log
for v in cur cb64 tile_16x8; do
OSSG_B200_LIB=$PWD/build/variants/libossg_$v.so timeout 120 python tools/slab_time.py 2>&1 | tail -1
done | tee gpurun_out/r02l_slabs.log
