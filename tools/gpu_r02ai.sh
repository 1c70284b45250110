#!/bin/bash
# round-2 GPU call AI (1 GPU): final verification at HEAD: whole GPU suite (plain and poisoned), smoke, bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02ai_pytest.log 2>&1
tail -6 gpurun_out/r02ai_pytest.log
( time OSSG_TEST_POISON=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02ai_pytest_poison.log 2>&1
tail -4 gpurun_out/r02ai_pytest_poison.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02ai_bench.json 2> gpurun_out/r02ai_bench.err
tail -c 200 gpurun_out/r02ai_bench.err
python tools/show_bench.py gpurun_out/r02ai_bench.json 2>/dev/null | head -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ai_bench_ref.json 2> gpurun_out/r02ai_bench_ref.err
head -c 300 gpurun_out/r02ai_bench_ref.json
