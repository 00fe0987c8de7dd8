#!/bin/bash
# Every GPU test, the inference bench (with per-launch breakdown) and the train bench at B=4 / B=32.   usage: tools/gpu_check.sh [steps]
mkdir -p gpurun_out
STEPS="${1:-100}"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
python tools/show_breakdown.py gpurun_out/breakdown.json
for B in 4 32; do
  timeout 300 python bench_train.py --steps 20 --warmup 3 --batch $B > gpurun_out/train_b$B.log 2>&1; echo "train B=$B rc=$?"; tail -1 gpurun_out/train_b$B.log | cut -c1-200
done
