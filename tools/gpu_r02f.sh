#!/bin/bash
# round-2 GPU call F (1 GPU): tests, variant timings (frame in the tile kernel's first blocks vs frame behind the tile kernel vs fork/join), bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02f_pytest.log 2>&1
tail -5 gpurun_out/r02f_pytest.log
for rep in 1 2; do
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done
done | tee gpurun_out/r02f_variants.log
timeout 600 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 300 gpurun_out/r02f_bench.err
python tools/show_bench.py gpurun_out/r02f_bench.json
