#!/bin/bash
# A/B of one environment switch on the same box.  usage: tools/gpu_ab_env.sh VAR "<value A>" "<value B>" [pytest -k expr]
mkdir -p gpurun_out
VAR="$1"; A="$2"; B="$3"; K="${4:-}"
if [ -n "$K" ]; then
  env $VAR="$B" timeout 900 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/t_ab.log 2>&1; echo "pytest($VAR=$B) rc=$?"; tail -4 gpurun_out/t_ab.log
fi
for V in "$A" "$B" "$A" "$B"; do
  T=$(echo "$V" | tr -c 'A-Za-z0-9_.-' '_')
  env $VAR="$V" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/breakdown_$T.json > gpurun_out/bench_ab_$T.log 2> gpurun_out/bench_ab_$T.err; echo "$VAR=$V rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_ab_$T.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
