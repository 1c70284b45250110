#!/bin/bash
# round-2 GPU call AE (1 GPU): north-halo coordinate rows through the straight-line point: tests (poisoned + plain), timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do timeout 120 python tools/variant_time.py 2>&1 | tail -1; timeout 120 python tools/slab_time.py 2>&1 | tail -1; done | tee gpurun_out/r02ae_times.log
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02ae_pytest_poison.log 2>&1
tail -4 gpurun_out/r02ae_pytest_poison.log
timeout 600 python bench.py > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err
python tools/show_bench.py gpurun_out/r02ae_bench.json 2>/dev/null | head -4
