#!/bin/bash
# Variant builds of the library that differ only in the BatchNorm accumulator kernels' tunables (csrc/train_bn.cu: BN_RED_U, BN_RED_MINB,
# BN_APPLY_U, BN_APPLY_MINB), for tools/diag_bn_kernels.py / tools/gpu_bn_variants.sh.  tools/bin/libhulk_bn<name>.so; "prev" = HEAD's file.
set -e
cd "$(dirname "$0")/.."
python -m hulk_keypoints_b200.build > /dev/null
B=hulk_keypoints_b200/csrc/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OTHERS=$(ls $B/*.o | grep -v "\.diag\.o" | grep -v train_bn)
mkdir -p tools/bin
build() {  # name, source, flags...
  local name=$1 src=$2; shift 2
  nvcc $FLAGS "$@" -Ihulk_keypoints_b200/csrc -c $src -o /tmp/train_bn_$name.o
  nvcc -shared -o tools/bin/libhulk_bn$name.so /tmp/train_bn_$name.o $OTHERS -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
  echo built tools/bin/libhulk_bn$name.so
}
S=hulk_keypoints_b200/csrc/train_bn.cu
build m0 $S &
build r1 $S -DBN_RED_MIN_CHUNKS=1 &
build r2 $S -DBN_RED_MIN_CHUNKS=2 &
build w2 $S -DBN_APPLY_WAVES=2 &
wait
build r1w2 $S -DBN_RED_MIN_CHUNKS=1 -DBN_APPLY_WAVES=2 &
build u2 $S -DBN_RED_MIN_CHUNKS=2 -DBN_APPLY_U=2 -DBN_APPLY_MINB=4 -DBN_APPLY_CONTIG=0 &
wait
