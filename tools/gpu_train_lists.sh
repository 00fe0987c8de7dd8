#!/bin/bash
# Fresh per-launch lists (duration + DRAM bytes) of one train step at B=4 and B=32 (eager launches; each ncu run follows a plain run).
mkdir -p gpurun_out
for B in 4 32; do
  CMD="python bench_train.py --steps 1 --warmup 3 --batch $B"
  HK_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/train_plain_b$B.log 2>&1 &&
  HK_TRAIN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches_b$B.csv $CMD > gpurun_out/ncu_train_b$B.log 2>&1
  echo "B=$B launch list rc=$?"
  python tools/train_list_summary.py gpurun_out/train_launches_b$B.csv --last-step gpurun_out/train_step_b${B}_launches.csv 2>&1 | head -32
  timeout 300 python bench_train.py --steps 20 --warmup 3 --batch $B > gpurun_out/train_b$B.log 2>&1; tail -1 gpurun_out/train_b$B.log | cut -c1-200
done
