#!/bin/bash
# Full GPU round: tests, smoke, bench (+breakdown), reference arm, train bench, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-2500; tail -3 gpurun_out/bench.err
timeout 300 python bench_train.py --steps 10 --warmup 3 > gpurun_out/train_n1.log 2>&1; echo "train rc=$?"; grep '^{' gpurun_out/train_n1.log | cut -c1-800
CMD="python bench.py --steps 2 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -s 60 -c 4 -o gpurun_out/prof_conv_tc2 $CMD > gpurun_out/ncu_conv2.log 2>&1
echo "conv_tc2 full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stem_tc_kernel|conv_tc_c64|head_|maxpool|argmax" -s 12 -c 12 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "small full rc=$?"
ls -la gpurun_out
