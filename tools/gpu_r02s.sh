#!/bin/bash
# round-2 GPU call S (1 GPU): fused fill with full 64-byte granule stores vs the shipped form, DRAM bytes of both
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do for g in 0 1; do OSSG_FILL_GRANULE=$g timeout 120 python tools/fill_time.py 2>&1 | tail -1; done; done | tee gpurun_out/r02s_fill.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fill or zipper" 2>&1 | tail -2
OSSG_FILL_GRANULE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fill or zipper" 2>&1 | tail -2
for g in 0 1; do
OSSG_FILL_GRANULE=$g timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:fill_fused --launch-skip 3 --launch-count 3 --csv --log-file gpurun_out/r02s_ncu_fill_g$g.csv python tools/prof_fill.py --fill > gpurun_out/r02s_ncu_g$g.log 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/r02s_ncu_fill_g$g.csv | cut -d, -f5,13- | head -12
done
