#!/bin/bash
# round-2 GPU call P (1 GPU): coordinate-kernel variants (unrolled y at 64 / 80 registers, 80 registers)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02p_variants.log
