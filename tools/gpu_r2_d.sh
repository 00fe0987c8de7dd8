#!/bin/bash
# BatchNorm accumulator path: kernel tests, train engine tests, train-step A/B (HK_BN_ACC=1/0) at per-GPU batch 4 and 32.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "bn_" > gpurun_out/t_bn.log 2>&1; echo "bn tests rc=$?"; tail -15 gpurun_out/t_bn.log
timeout 1200 python -m pytest tests/test_gpu_train_engine.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train engine tests rc=$?"; tail -8 gpurun_out/t_train.log
for acc in 1 0 1 0; do
  for B in 4 32; do
    HK_BN_ACC=$acc timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_acc${acc}_b$B.log 2>&1; echo "acc=$acc B=$B rc=$? $(tail -1 gpurun_out/train_acc${acc}_b$B.log | cut -c1-260)"
  done
done
