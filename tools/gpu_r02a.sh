#!/bin/bash
# round-2 GPU call A (1 GPU): tests, fill variants, ncu DRAM captures of the fill kernels, memcheck, bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02a_pytest.log 2>&1
timeout 400 python tools/fill_variants.py > gpurun_out/r02a_fill_variants.jsonl 2> gpurun_out/r02a_fill_variants.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:zipper_kernel --launch-skip 20 --launch-count 4 --csv --log-file gpurun_out/r02a_ncu_zipper_loop8.csv python tools/prof_fill.py --loop8 > gpurun_out/r02a_ncu1.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:zipper_kernel --launch-skip 3 --launch-count 4 --csv --log-file gpurun_out/r02a_ncu_zipper_flush4.csv python tools/prof_fill.py --zipper4 > gpurun_out/r02a_ncu2.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:fill_fused --launch-skip 3 --launch-count 4 --csv --log-file gpurun_out/r02a_ncu_fill_flush4.csv python tools/prof_fill.py --fill > gpurun_out/r02a_ncu3.log 2>&1
OSSG_PEER_TIMEOUT_MS=20000 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_case.py > gpurun_out/r02a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02a_memcheck.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 400 gpurun_out/r02a_bench.err
cat gpurun_out/r02a_pytest.log | tail -8
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done > gpurun_out/r02a_variants.log
cat gpurun_out/r02a_variants.log
