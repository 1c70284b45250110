#!/bin/bash
# round-2 GPU call B (1 GPU): tests, fill kernels' DRAM captures with the branch-free sign, L2::64B hint, bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02b_pytest.log 2>&1
timeout 400 python tools/fill_variants.py > gpurun_out/r02b_fill_variants.jsonl 2> gpurun_out/r02b_fill_variants.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:zipper_kernel --launch-skip 20 --launch-count 4 --csv --log-file gpurun_out/r02b_ncu_zipper_loop8.csv python tools/prof_fill.py --loop8 > gpurun_out/r02b_ncu1.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:zipper_kernel --launch-skip 3 --launch-count 4 --csv --log-file gpurun_out/r02b_ncu_zipper_flush4.csv python tools/prof_fill.py --zipper4 > gpurun_out/r02b_ncu2.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:fill_fused --launch-skip 3 --launch-count 4 --csv --log-file gpurun_out/r02b_ncu_fill_flush4.csv python tools/prof_fill.py --fill > gpurun_out/r02b_ncu3.log 2>&1
OSSG_FILL_VARIANT=31 timeout 300 ncu --cache-control none --clock-control none --metrics $M -k regex:fill_fused --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/r02b_ncu_fill_flush4_v31.csv python tools/prof_fill.py --fill > gpurun_out/r02b_ncu4.log 2>&1
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -c 400 gpurun_out/r02b_bench.err
tail -8 gpurun_out/r02b_pytest.log
