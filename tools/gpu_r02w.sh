#!/bin/bash
# round-2 GPU call W (1 GPU): whole GPU test suite at HEAD (incl. the launch-option child processes), smoke, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02w_pytest.log 2>&1
tail -6 gpurun_out/r02w_pytest.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
tail -c 200 gpurun_out/r02w_bench.err
python tools/show_bench.py gpurun_out/r02w_bench.json | head -4
