#!/bin/bash
# BatchNorm accumulator kernels: stand-alone timing of every tools/bin/libhulk_bn*.so variant (tools/build_bn_variants.sh) with CUDA events and,
# per kernel, under ncu; the BatchNorm kernel tests on the in-tree build; the train step (batch 32 and 4) with the previous and the new kernels.
mkdir -p gpurun_out
VARS=${VARS:-"prev a b c d e f"}
TRAIN_VARS=${TRAIN_VARS:-"prev a c"}
for V in $VARS; do
  HK_LIB_PATH=$PWD/tools/bin/libhulk_bn$V.so timeout 300 python tools/diag_bn_kernels.py --iters 10 --json gpurun_out/bn_$V.json > gpurun_out/bn_$V.log 2>&1
  echo "variant $V rc=$? $(tail -1 gpurun_out/bn_$V.log)"
done
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -x -q -m gpu -k "bn" > gpurun_out/t_bn.log 2>&1; echo "bn tests rc=$? $(tail -1 gpurun_out/t_bn.log)"
for rep in 1 2; do
  for V in $TRAIN_VARS; do
    for B in 32 4; do
      HK_LIB_PATH=$PWD/tools/bin/libhulk_bn$V.so timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_bn${V}_b$B.log 2>&1
      echo "lib=$V B=$B rc=$? $(tail -1 gpurun_out/train_bn${V}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
for V in $VARS; do
  HK_LIB_PATH=$PWD/tools/bin/libhulk_bn$V.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/bn_${V}_launches.csv python tools/diag_bn_kernels.py --iters 1 --shapes layer1,layer3,layer4 > gpurun_out/bn_${V}_ncu.log 2>&1; echo "ncu $V rc=$?"
done
