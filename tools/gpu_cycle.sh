python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['kernel_ms'],d['zipper'])"
