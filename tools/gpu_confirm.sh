#!/bin/bash
# Full confirmation on a fresh box: every GPU test, smoke, the default bench (+ breakdown), the reference arm, the train bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
python tools/show_breakdown.py gpurun_out/breakdown.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
for B in 4 32; do
  timeout 300 python bench_train.py --steps 20 --warmup 3 --batch $B > gpurun_out/train_b$B.log 2>&1; echo "train B=$B rc=$?"; tail -1 gpurun_out/train_b$B.log | cut -c1-200
done
