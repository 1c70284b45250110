timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_fold or zipper" 2>&1 | tail -5
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['kernel_ms']);print(json.dumps(d['zipper'],indent=0)[:2500]);print(d['roofline_zipper']);print(d['cpu_baseline']);print(d['e2e'])"
