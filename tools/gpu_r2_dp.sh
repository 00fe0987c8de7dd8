#!/bin/bash
# Data-parallel train step on N GPUs: bit-equality check against the local step (tools/check_dp_step.py), then the train bench.
# (HK_DP_SPLIT_ADAM was the A/B switch of an experiment that is no longer in the tree: the late Adam update beside the early backward.)
# usage (gpurun --gpus N): tools/gpu_r2_dp.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/check_dp_step.py > gpurun_out/dp_check.log 2>&1; echo "dp check rc=$?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/dp_check.log | tail -4
for rep in 1 2; do
  for S in 1 0; do
    for B in ${BATCHES:-4}; do
      HK_DP_SPLIT_ADAM=$S timeout 300 $TR --master-port 2954$S bench_train.py --gpus $N --steps 40 --warmup 5 --batch $B > gpurun_out/dp_split${S}_b${B}_$rep.log 2>&1
      echo "split=$S N=$N B=$B rc=$? $(grep '^{' gpurun_out/dp_split${S}_b${B}_$rep.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
