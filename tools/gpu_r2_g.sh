#!/bin/bash
# Head fusion (last conv + K scoring rows): kernel + model tests under a watchdog, then A/B HK_FUSE_HEAD=1/0 on one box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_head" > gpurun_out/t_head.log 2>&1; echo "conv_head tests rc=$?"; tail -15 gpurun_out/t_head.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -6 gpurun_out/t_model.log
for rep in 1 2; do
for mode in 1 0; do
  HK_FUSE_HEAD=$mode timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-config3 --no-sustained --breakdown gpurun_out/breakdown_h$mode.json > gpurun_out/bench_h$mode.log 2> gpurun_out/bench_h$mode.err; echo "bench HK_FUSE_HEAD=$mode rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_h$mode.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches_per_step"], "clk", d["clocks"]["sm_mhz"])
b = json.load(open("gpurun_out/breakdown_h$mode.json"))
print("  ", [(r["name"], round(r["ms"], 3)) for r in b["rows"] if r["name"].startswith(("b15.", "head", "upsample", "decode"))])
PY
done
done
