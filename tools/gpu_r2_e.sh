#!/bin/bash
# Conv-epilogue BatchNorm statistics: kernel tests (under a watchdog), train engine tests, train-step A/B (HK_CONV_STATS=1/0) at per-GPU batch 4 and 32.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "conv_epilogue or bn_" > gpurun_out/t_cs.log 2>&1; echo "conv stats tests rc=$?"; tail -15 gpurun_out/t_cs.log
timeout 1200 python -m pytest tests/test_gpu_train_engine.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train engine tests rc=$?"; tail -8 gpurun_out/t_train.log
for cs in 1 0 1 0; do
  for B in 4 32; do
    HK_CONV_STATS=$cs timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_cs${cs}_b$B.log 2>&1; echo "conv_stats=$cs B=$B rc=$? $(tail -1 gpurun_out/train_cs${cs}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
  done
done
