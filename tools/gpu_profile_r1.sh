#!/bin/bash
# Round-1 ncu evidence for the final inference kernels (GPU box).  Each ncu command runs only after the identical plain
# command exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"conv_tc|stem_tc|head_|maxpool|argmax" -s 123 -c 41 --csv --log-file gpurun_out/step_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "per-launch dram rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_tc2h_kernel" -s 96 -c 10 -o gpurun_out/prof_conv_tc2 $CMD > gpurun_out/ncu_conv2.log 2>&1
echo "conv_tc2 full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stem_tc_kernel|conv_tc_c64|head_|maxpool|argmax" -s 36 -c 12 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "small full rc=$?"
ls -la gpurun_out | head -30
