#!/bin/bash
# Runs on the GPU box (via gpurun): every stage under its own timeout, logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== kernels (no tcgen05) ==" 
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "not tcgen05" > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_kernels.log
echo "== diag tcgen05 =="
timeout 300 python tools/diag_conv_tc.py > gpurun_out/diag_tc.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/diag_tc.log
echo "== tcgen05 tests =="
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_tc.log
echo "== smoke =="
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/smoke.log
echo "== model tests =="
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_model.log
echo "== bench =="
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/breakdown.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
