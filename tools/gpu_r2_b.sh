#!/bin/bash
# stem+pool fusion check: new kernel tests (under a watchdog), the model tests, a short bench with and without the fused kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stem_pool" > gpurun_out/t_stempool.log 2>&1; echo "stem_pool tests rc=$?"; tail -15 gpurun_out/t_stempool.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -6 gpurun_out/t_model.log
for mode in 1 0; do
HK_STEM_POOL=$mode timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-config3 --no-sustained --breakdown gpurun_out/breakdown_sp$mode.json > gpurun_out/bench_sp$mode.log 2> gpurun_out/bench_sp$mode.err; echo "bench HK_STEM_POOL=$mode rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_sp$mode.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches_per_step"], "clk", d["clocks"]["sm_mhz"], [(r["kernel"], round(r["ms"],3), round(r["frac"],2)) for r in d["roofline_hbm_kernels"]])
PY
done
