#!/bin/bash
# ncu evidence (GPU box).  Each ncu command runs only after the identical plain command exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 133 -c 6 -o gpurun_out/prof_conv_tc $CMD > gpurun_out/ncu_conv.log 2>&1
echo "conv_tc full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stem_tc_kernel|head_|maxpool|argmax" -s 5 -c 6 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "small full rc=$?"
tail -3 gpurun_out/plain.log
ls -la gpurun_out
