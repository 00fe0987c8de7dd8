#!/bin/bash
# usage: tools/gpu_quick.sh "<pytest -k expression or empty for all>" [bench steps]
mkdir -p gpurun_out
K="$1"; STEPS="${2:-100}"
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_quick.log
else
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_all.log
fi
timeout 600 python bench.py --steps $STEPS --warmup 5 --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-1500; tail -5 gpurun_out/bench.err
