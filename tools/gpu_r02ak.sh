#!/bin/bash
# round-2 GPU call AK (1 GPU): bench with the graph-replay keys of the small configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02ak_bench.json 2> gpurun_out/r02ak_bench.err
tail -c 300 gpurun_out/r02ak_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02ak_bench.json'))
print(d['ms_per_step'], d['parity']['mismatches'])
for k,v in d['configs'].items(): print(k, v['construct_f64'])
PY
