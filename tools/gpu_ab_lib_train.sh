#!/bin/bash
# A/B of two builds of the library on the train step (same box): tools/bin/libhulk_prev.so vs tools/bin/libhulk_new.so, alternating.
mkdir -p gpurun_out
for rep in 1 2; do
  for L in prev new; do
    for B in 4 32; do
      HK_LIB_PATH=$PWD/tools/bin/libhulk_$L.so timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_lib_${L}_b$B.log 2>&1
      echo "lib=$L B=$B rc=$? $(tail -1 gpurun_out/train_lib_${L}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
