#!/bin/bash
# Round 2, first GPU pass: pin the trained fixtures, run the new parity tests verbosely, then every GPU test, smoke() and a short bench.
mkdir -p gpurun_out
timeout 600 python tools/pin_ftrn.py > gpurun_out/pin.log 2>&1; echo "pin rc=$?"; cat gpurun_out/pin.log | tail -5
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -rA > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -25 gpurun_out/t_parity.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
