#!/bin/bash
# stem_pool epilogue on raw accumulators: bit-identity tests, model tests, then inference A/B of tools/bin/libhulk_prev.so / libhulk_new.so.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stem" > gpurun_out/t_sp.log 2>&1; echo "stem tests rc=$?"; tail -8 gpurun_out/t_sp.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -5 gpurun_out/t_model.log
for rep in 1 2; do
for L in prev new; do
  HK_LIB_PATH=$PWD/tools/bin/libhulk_$L.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-config3 --no-sustained --breakdown gpurun_out/breakdown_$L.json > gpurun_out/bench_$L.log 2> gpurun_out/bench_$L.err; echo "bench $L rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_$L.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "e2e_u8", round(d["e2e_uint8_input"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"])
b = json.load(open("gpurun_out/breakdown_$L.json"))
print("  ", [(r["name"], round(r["ms"], 3)) for r in b["rows"] if r["name"].startswith(("stem", "upsample", "decode"))])
PY
done
done
