#!/bin/bash
# round-2 GPU call U (1 GPU): periodic part of the fused fill at constant row count and varying row pitch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/fill_pitch_probe.py 2>&1 | tee gpurun_out/r02u_fill_pitch.log
