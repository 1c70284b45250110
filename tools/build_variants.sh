#!/bin/bash
# Build one libossg_b200 per kernel experiment into build/variants/ (git-ignored, travels with gpurun);
# `bash tools/gpu_variants.sh` on the GPU box then times each on the 1/12 degree construction and prints an
# output checksum -- identical checksums = identical bits, so a faster variant can be adopted directly.
#   cur             the shipped configuration (working tree, Makefile flags)
#   head            the last commit's sources (git worktree), same flags: the baseline of the working tree's changes
#   fprange         -DOSSG_RANGE_FP_TESTS     tile kernel: FP64 comparisons for the speculative range tests (round-1 form)
#   lonsel          -DOSSG_LON_SPEC_SELECTS   tile kernel: select-chain form of sincosd_lon_spec (round-1 form)
#   coordunr        -DOSSG_COORD_UNROLL       coordinate kernel: the four staggerings unrolled (compile-time locations)
#   cxrolled        -DOSSG_COORD_SPLIT_X=0    coordinate kernel: one rolled loop over the four staggerings (default: one copy per x location)
#   cxy64, cxy80    -DOSSG_COORD_UNROLL_Y [-DOSSG_COORD_MINBLOCKS=6]  all four staggerings unrolled, at 64 / 80 registers; cx80: default form at 80 registers
#   pcpair          -DOSSG_PC_SOA=0           tile kernel: deg2rad(phi)/2 and cos(phi) of a patch point as one 16-byte shared element
#   cb256, cb64     -DOSSG_COORD_BLOCK=..     coordinate kernel: 256- / 64-thread blocks (default 128, 8 per SM)
# (fprange, lonsel, coordunr only with --all)
#   tile_TXxTY      -DOSSG_TX=.. -DOSSG_TY=.. other tile shapes of the metric kernel (measured in round 1: 32x8 stays)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=$ROOT/build/variants
mkdir -p $OUT
rm -f $OUT/*.so
BASE="-DOSSG_ATAN_RANGE_TABLE -DOSSG_COORD_MAGIC_DIV"
build() { # srcdir name flags...
  local src=$1 name=$2; shift 2
  ( cd $src && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC $BASE "$@" -shared \
       -o $OUT/libossg_$name.so ossg_api.cu ossg_tripolar.cu ossg_zipper.cu ossg_peer.cu ossg_nccl.cu -cudart shared -ldl )
  echo "built $name"
}
CS=$ROOT/orthogonalsphericalshellgrids.jl_b200/csrc
build $CS cur &
if [ "$1" == "--all" ]; then
  shift
  build $CS fprange -DOSSG_RANGE_FP_TESTS &
  build $CS lonsel -DOSSG_LON_SPEC_SELECTS &
  build $CS coordunr -DOSSG_COORD_UNROLL &
fi
build $CS pcpair -DOSSG_PC_SOA=0 &
build $CS cxrolled -DOSSG_COORD_SPLIT_X=0 &
build $CS cxy64 -DOSSG_COORD_UNROLL_Y &
build $CS cxy80 -DOSSG_COORD_UNROLL_Y -DOSSG_COORD_MINBLOCKS=6 &
build $CS cx80 -DOSSG_COORD_MINBLOCKS=6 &
build $CS cb256 -DOSSG_COORD_BLOCK=256 &
build $CS cb64 -DOSSG_COORD_BLOCK=64 &
wait
if [ "$1" == "--head" ]; then
  shift
  rm -rf /tmp/ossg_head && git -C $ROOT worktree prune && git -C $ROOT worktree add -f /tmp/ossg_head HEAD >/dev/null 2>&1
  build /tmp/ossg_head/orthogonalsphericalshellgrids.jl_b200/csrc head
  git -C $ROOT worktree remove --force /tmp/ossg_head
fi
for shape in "$@"; do build $CS tile_$shape -DOSSG_TX=${shape%x*} -DOSSG_TY=${shape#*x}; done
