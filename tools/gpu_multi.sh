#!/bin/bash
# N-GPU checks (gpurun --gpus N): sharded inference bench + data-parallel train bench (TrainEngine + NCCL all-reduce of the flat gradient).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; grep '^{' gpurun_out/bench_n$N.log | cut -c1-330; tail -3 gpurun_out/bench_n$N.err
for B in 4 32; do
  timeout 600 $TR --master-port 29518 bench_train.py --gpus $N --steps 20 --warmup 3 --batch $B > gpurun_out/train_n${N}_b$B.log 2> gpurun_out/train_n${N}_b$B.err; echo "train N=$N B=$B rc=$?"; grep '^{' gpurun_out/train_n${N}_b$B.log | cut -c1-260; tail -3 gpurun_out/train_n${N}_b$B.err
done
timeout 300 python bench_train.py --steps 20 --warmup 3 > gpurun_out/train_n1.log 2>&1; echo "train N=1 rc=$?"; grep '^{' gpurun_out/train_n1.log | cut -c1-260
