#!/bin/bash
mkdir -p gpurun_out
echo "== stem test =="
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stem" > gpurun_out/t_stem.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_stem.log
echo "== model tests =="
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_model.log
echo "== bench =="
timeout 600 python bench.py --steps 100 --warmup 5 --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
