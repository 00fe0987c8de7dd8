#!/bin/bash
# Block-entry fusion check: kernel tests (under a watchdog), the model tests, short benches: fused / fused without early release / unfused.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_ds" > gpurun_out/t_ds.log 2>&1; echo "conv_ds tests rc=$?"; tail -15 gpurun_out/t_ds.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -6 gpurun_out/t_model.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-config3 --no-sustained --breakdown gpurun_out/breakdown_$name.json > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_$name.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches_per_step"], "clk", d["clocks"]["sm_mhz"])
b = json.load(open("gpurun_out/breakdown_$name.json"))
print("  ", [(r["name"], round(r["ms"], 3)) for r in b["rows"] if r["name"].startswith(("b3.", "b7.", "b13."))])
PY
}
run ds1 HK_FUSE_DS=1
run ds0 HK_FUSE_DS=0
run ds1_late HK_FUSE_DS=1 HK_DS_EARLY=0
run ds1b HK_FUSE_DS=1
run ds0b HK_FUSE_DS=0
