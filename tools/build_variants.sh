#!/bin/bash
# Build one libossg_b200 per kernel experiment into build/variants/ (git-ignored, travels with gpurun);
# `bash tools/gpu_variants.sh` on the GPU box then times each on the 1/12 degree construction and prints an
# output checksum -- identical checksums = identical bits, so a faster variant can be adopted directly.
#   base            the shipped configuration
#   atanrng         -DOSSG_ATAN_RANGE_TABLE   coordinate kernel: atan range from a shared-memory byte table (hot loop 235 -> 223 instructions)
#   magicdiv        -DOSSG_COORD_MAGIC_DIV    coordinate kernel: multiply-shift instead of the integer division of the block map
#   coordx          both of the above
#   tile_TXxTY      -DOSSG_TX=.. -DOSSG_TY=.. other tile shapes of the metric kernel (measured in round 1: 32x8 stays)
set -e
cd "$(dirname "$0")/../orthogonalsphericalshellgrids.jl_b200/csrc"
OUT=../../build/variants
mkdir -p $OUT
build() { # name flags...
  local name=$1; shift
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -shared \
       -o $OUT/libossg_$name.so ossg_api.cu ossg_tripolar.cu ossg_zipper.cu ossg_peer.cu ossg_nccl.cu -cudart shared -ldl
  echo "built $name"
}
build base            # (the plain forms: no atan range table, integer division in the block map)
build atanrng -DOSSG_ATAN_RANGE_TABLE
build magicdiv -DOSSG_COORD_MAGIC_DIV
build coordx -DOSSG_ATAN_RANGE_TABLE -DOSSG_COORD_MAGIC_DIV
for shape in "$@"; do build tile_$shape -DOSSG_TX=${shape%x*} -DOSSG_TY=${shape#*x}; done
