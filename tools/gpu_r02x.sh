#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/slab_profile.py 2>&1 | tee gpurun_out/r02x_slab_profile.log
