#!/bin/bash
# Separable upsample backward: head tests, train engine tests, train-step A/B (HK_UPSAMPLE_BWD_ROWS=1/0).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "head" > gpurun_out/t_hb.log 2>&1; echo "head tests rc=$?"; tail -8 gpurun_out/t_hb.log
timeout 1200 python -m pytest tests/test_gpu_train_engine.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train engine tests rc=$?"; tail -6 gpurun_out/t_train.log
for v in 1 0 1 0; do
  for B in 4 32; do
    HK_UPSAMPLE_BWD_ROWS=$v timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_ub${v}_b$B.log 2>&1; echo "rows=$v B=$B rc=$? $(tail -1 gpurun_out/train_ub${v}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
  done
done
