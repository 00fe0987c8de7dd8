#!/bin/bash
# per-kernel times of one training step (eager launches), B given as $1 (default 4)
mkdir -p gpurun_out
B=${1:-4}
CMD="python bench_train.py --steps 2 --warmup 3 --batch $B"
HK_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/train_plain.log 2>&1 &&
HK_TRAIN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/train_launches_b$B.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/train_plain.log | cut -c1-300
