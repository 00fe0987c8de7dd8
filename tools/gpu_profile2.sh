#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stem_tc_kernel|head_|maxpool" -s 4 -c 4 -o gpurun_out/prof_small2 $CMD > gpurun_out/ncu_small2.log 2>&1
echo "small rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel<64>|conv_tc_kernel<128>" -s 24 -c 4 -o gpurun_out/prof_l12 $CMD > gpurun_out/ncu_l12.log 2>&1
echo "l12 rc=$?"
