#!/bin/bash
# Full confirmation on a fresh box: every GPU test, smoke, the default bench, the train bench, and a fresh train-step launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
for B in 4 32; do
  timeout 300 python bench_train.py --steps 20 --warmup 3 --batch $B > gpurun_out/train_b$B.log 2>&1; echo "train B=$B rc=$?"; tail -1 gpurun_out/train_b$B.log | cut -c1-300
done
CMD="python bench_train.py --steps 1 --warmup 3 --batch 32"
HK_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/train_plain.log 2>&1 &&
HK_TRAIN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches_b32.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo "train launch list rc=$?"
