#!/bin/bash
# Tap-pair weight gradient of the 64-channel layer: tests, then train-step A/B (HK_WGRAD_TAP_PAIRS=1/0) at per-GPU batch 4 and 32.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/t_wg.log 2>&1; echo "wgrad tests rc=$?"; tail -12 gpurun_out/t_wg.log
timeout 1200 python -m pytest tests/test_gpu_train_engine.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train engine tests rc=$?"; tail -6 gpurun_out/t_train.log
for tp in 1 0 1 0; do
  for B in 4 32; do
    HK_WGRAD_TAP_PAIRS=$tp timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_tp${tp}_b$B.log 2>&1; echo "tap_pairs=$tp B=$B rc=$? $(tail -1 gpurun_out/train_tp${tp}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
  done
done
