#!/bin/bash
# 2-GPU pass: inference bench (no collective) + the train step with its NCCL share.
mkdir -p gpurun_out
N="${1:-2}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_n$N.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sustained", round(d["sustained"]["value"]), "config3", d["config3_4096_images"]["value"])
        print(json.dumps(d["train_step"], indent=1))
PY
tail -3 gpurun_out/bench_n$N.err
