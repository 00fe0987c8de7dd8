#!/bin/bash
# Max-pool forward with indices in the train step: tests + train-step A/B (HK_MAXPOOL_IDX=1/0).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "maxpool" > gpurun_out/t_mp.log 2>&1; echo "maxpool tests rc=$?"; tail -8 gpurun_out/t_mp.log
timeout 1200 python -m pytest tests/test_gpu_train_engine.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train engine tests rc=$?"; tail -6 gpurun_out/t_train.log
for v in 1 0 1 0; do for B in 4 32; do
  HK_MAXPOOL_IDX=$v timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_mp${v}_b$B.log 2>&1
  echo "maxpool_idx=$v B=$B rc=$? $(tail -1 gpurun_out/train_mp${v}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
done; done
