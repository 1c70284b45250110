timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_final.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a gpurun_out/pytest_final.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['kernel_ms']);print(d['roofline_zipper']);print(d['cpu_baseline']);print(d['e2e']['value'])"
