#!/bin/bash
# Cheap part of tools/gpu_profile_r1.sh: launch list + per-launch DRAM / tensor-pipe metrics of one inference step.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"conv_tc|stem_tc|head_|maxpool|argmax" -s 123 -c 41 --csv --log-file gpurun_out/step_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "per-launch dram rc=$?"
