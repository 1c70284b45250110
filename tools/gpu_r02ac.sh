#!/bin/bash
# round-2 GPU call AC (1 GPU): tests, bench (own arm and reference arm), launch list of the bench, full ncu capture of the construction kernels at HEAD
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02ac_pytest.log 2>&1
tail -4 gpurun_out/r02ac_pytest.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02ac_bench.json 2> gpurun_out/r02ac_bench.err
tail -c 300 gpurun_out/r02ac_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ac_bench_ref.json 2> gpurun_out/r02ac_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ac_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02ac_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'metrics_tile|coords_pair|tables_kernel|metrics_frame' --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r02ac python tools/prof_step.py > gpurun_out/r02ac_ncu_full.log 2>&1
tail -3 gpurun_out/r02ac_ncu_full.log
python tools/show_bench.py gpurun_out/r02ac_bench.json
for rep in 1 2; do timeout 120 python tools/variant_time.py 2>&1 | tail -1; timeout 120 python tools/slab_time.py 2>&1 | tail -1; done | tee gpurun_out/r02ac_times.log
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02ac_pytest_poison.log 2>&1
tail -4 gpurun_out/r02ac_pytest_poison.log
