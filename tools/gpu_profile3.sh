#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-8} -c ${3:-3} -o gpurun_out/prof_$4 $CMD > gpurun_out/ncu_$4.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_$4.log
