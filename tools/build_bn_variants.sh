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
git show HEAD:hulk_keypoints_b200/csrc/train_bn.cu > /tmp/train_bn_prev.cu
build prev /tmp/train_bn_prev.cu &
S=hulk_keypoints_b200/csrc/train_bn.cu
build a $S -DBN_APPLY_U=4 -DBN_APPLY_MINB=3 -DBN_APPLY_CONTIG=1 &
build b $S -DBN_APPLY_U=4 -DBN_APPLY_MINB=3 -DBN_APPLY_CONTIG=0 &
build c $S -DBN_APPLY_U=2 -DBN_APPLY_MINB=4 -DBN_APPLY_CONTIG=0 &
wait
build d $S -DBN_APPLY_U=8 -DBN_APPLY_MINB=2 -DBN_APPLY_CONTIG=1 &
build e $S -DBN_APPLY_U=4 -DBN_APPLY_MINB=4 -DBN_APPLY_CONTIG=1 &
build f $S -DBN_APPLY_U=3 -DBN_APPLY_MINB=3 -DBN_APPLY_CONTIG=0 -DBN_BWDRED_U=3 &
wait
