#!/bin/bash
# Round-2 late: (1) A/B of the weight-gradient K-split policy (HK_WGRAD_ROUNDS=2 vs 1) on the train step, (2) with the winner exported:
# re-pin the trained fixtures (the BatchNorm / weight-gradient rounding changed), every GPU test, smoke.
mkdir -p gpurun_out
python -m hulk_keypoints_b200.build > /dev/null
for rep in 1 2; do
  for R in 2 1; do
    for B in 32 4; do
      HK_WGRAD_ROUNDS=$R timeout 300 python bench_train.py --steps 30 --warmup 5 --batch $B > gpurun_out/train_rounds${R}_b${B}_$rep.log 2>&1
      echo "rounds=$R B=$B rc=$? $(tail -1 gpurun_out/train_rounds${R}_b${B}_$rep.log | grep -o '"ms_per_step": [0-9.]*')"
    done
  done
done
WIN=$(python - <<'PY'
import json, glob
def ms(r, b):
    v = []
    for f in glob.glob(f"gpurun_out/train_rounds{r}_b{b}_*.log"):
        for l in open(f):
            if l.startswith("{"):
                v.append(json.loads(l)["ms_per_step"])
    return sum(v) / max(1, len(v))
gain = sum(ms(2, b) / ms(1, b) for b in (32, 4)) / 2    # > 1: one round is faster
print(1 if gain > 1.003 else 2)
PY
)
echo "HK_WGRAD_ROUNDS winner: $WIN"
export HK_WGRAD_ROUNDS=$WIN
timeout 900 python tools/pin_ftrn.py > gpurun_out/pin.log 2>&1; echo "pin rc=$?"; tail -4 gpurun_out/pin.log
cp tests/golden/ftrn_v2.json gpurun_out/ftrn_v2.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
