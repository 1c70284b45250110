#!/bin/bash
# round-2 GPU call E (1 GPU): tests with the working tree's library, variant timings (build/variants), public constructor profile
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02e_pytest.log 2>&1
tail -5 gpurun_out/r02e_pytest.log
for rep in 1 2; do
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02e_variants.log
timeout 300 python tools/public_ctor_profile.py > gpurun_out/r02e_public_ctor.log 2>&1
head -5 gpurun_out/r02e_public_ctor.log
