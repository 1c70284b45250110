# GPU tests + smoke with the default library, then time every library under build/variants/
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
for f in build/variants/*.so; do OSSG_B200_LIB=$PWD/$f timeout 120 python tools/variant_time.py 2>&1 | tail -1; done | tee gpurun_out/variants.log
