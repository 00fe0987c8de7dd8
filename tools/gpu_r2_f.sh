#!/bin/bash
# Single-box layer-1 kernel (16x8 patches, all nine taps from one haloed box): tests under a watchdog, then inference + train A/B of the
# two library builds tools/bin/libhulk_prev.so / libhulk_new.so on the same box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py -m gpu -q -x -k "c64 or conv_epilogue or conv_tcgen05_vs_oracle or dgrad" > gpurun_out/t_c64.log 2>&1; echo "c64 tests rc=$?"; tail -12 gpurun_out/t_c64.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -5 gpurun_out/t_model.log
for rep in 1 2; do
for L in prev new; do
  HK_LIB_PATH=$PWD/tools/bin/libhulk_$L.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-config3 --no-sustained --breakdown gpurun_out/breakdown_$L.json > gpurun_out/bench_$L.log 2> gpurun_out/bench_$L.err; echo "bench $L rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_$L.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"])
b = json.load(open("gpurun_out/breakdown_$L.json"))
print("  ", [(r["name"], round(r["ms"], 3)) for r in b["rows"] if r["name"].startswith(("b0.", "b1.", "b2."))])
PY
done
done
for L in prev new; do for B in 4 32; do
  HK_LIB_PATH=$PWD/tools/bin/libhulk_$L.so timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_lib_${L}_b$B.log 2>&1
  echo "train lib=$L B=$B rc=$? $(tail -1 gpurun_out/train_lib_${L}_b$B.log | grep -o '"ms_per_step": [0-9.]*')"
done; done
