#!/bin/bash
# A/B of several builds of the library on the train step (same box, alternating, two rounds).
# usage: tools/gpu_ab_libs_train.sh "<name> <name> ..." [batches]     (tools/bin/libhulk_<name>.so)
mkdir -p gpurun_out
NAMES="$1"; BATCHES="${2:-4 32}"
for rep in 1 2; do
  for L in $NAMES; do
    for B in $BATCHES; do
      HK_LIB_PATH=$PWD/tools/bin/libhulk_$L.so timeout 300 python bench_train.py --steps 40 --warmup 5 --batch $B > gpurun_out/train_lib_${L}_b${B}_$rep.log 2>&1
      echo "lib=$L B=$B rc=$? $(tail -1 gpurun_out/train_lib_${L}_b${B}_$rep.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
