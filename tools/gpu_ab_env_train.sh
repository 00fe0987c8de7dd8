#!/bin/bash
# A/B of one environment switch on the train step (same box, alternating, two rounds).  usage: tools/gpu_ab_env_train.sh VAR "<A>" "<B>" [batches]
mkdir -p gpurun_out
VAR="$1"; A="$2"; Bv="$3"; BATCHES="${4:-4 32}"
for rep in 1 2; do
  for V in "$A" "$Bv"; do
    for B in $BATCHES; do
      env $VAR="$V" timeout 300 python bench_train.py --steps 40 --warmup 5 --batch $B > gpurun_out/train_env_${V}_b${B}_$rep.log 2>&1
      echo "$VAR=$V B=$B rc=$? $(tail -1 gpurun_out/train_env_${V}_b${B}_$rep.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
