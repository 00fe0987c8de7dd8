#!/bin/bash
# N-GPU checks (gpurun --gpus N): sharded inference bench + DDP train bench.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; grep '^{' gpurun_out/bench_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench_train.py --gpus $N --steps 10 --warmup 3 > gpurun_out/train_n$N.log 2> gpurun_out/train_n$N.err; echo "train N=$N rc=$?"; grep '^{' gpurun_out/train_n$N.log | cut -c1-500; tail -3 gpurun_out/train_n$N.err
timeout 300 python bench_train.py --steps 10 --warmup 3 > gpurun_out/train_n1.log 2>&1; echo "train N=1 rc=$?"; grep '^{' gpurun_out/train_n1.log | cut -c1-400
timeout 300 python bench_train.py --steps 10 --warmup 3 --reference-loss > gpurun_out/train_n1_ref.log 2>&1; echo "train N=1 refloss rc=$?"; grep '^{' gpurun_out/train_n1_ref.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref arm rc=$?"; grep '^{' gpurun_out/bench_ref.log | cut -c1-600
