#!/bin/bash
# round-2 GPU call AJ (1 GPU): the constructor captured into a CUDA graph
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cuda_graph" 2>&1 | tail -15 | tee gpurun_out/r02aj_graph.log
