#!/bin/bash
# Data-parallel train step with the gradients exchanged as bf16 (HK_DP_BF16_GRADS=1, opt-in) against the fp32 exchange.  usage (gpurun --gpus N): tools/gpu_dp_bf16.sh N [batch]
N=${1:-2}; B=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for rep in 1 2; do
  for S in 0 1; do
    HK_DP_BF16_GRADS=$S timeout 300 $TR --master-port 2956$S bench_train.py --gpus $N --steps 40 --warmup 5 --batch $B > gpurun_out/dp_bf16_${S}_$rep.log 2>&1
    echo "bf16_grads=$S N=$N B=$B rc=$? $(grep '^{' gpurun_out/dp_bf16_${S}_$rep.log | grep -o '"ms_per_step": [0-9.]*\|"loss": [0-9.]*' | tr '\n' ' ')"
  done
done
