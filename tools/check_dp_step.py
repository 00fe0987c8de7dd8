"""Data-parallel train step == local train step, bit for bit, when every rank sees the same batch (the average of identical gradients
is the gradient itself for 2 / 4 / 8 ranks).  Exercises the two-graph backward, the averaging all-reduce of both ranges and the
range-by-range Adam update of train_ops.train_step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_dp_step.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_keypoints_b200 as hk  # noqa: E402
from hulk_keypoints_b200 import synth, train_ops  # noqa: E402
from hulk_keypoints_b200.optim import FusedAdam  # noqa: E402


def run(exchange: bool, steps: int = 4):
    torch.manual_seed(0)
    m = hk.KeypointsGauss(4).cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator().manual_seed(7)
    losses = []
    for _ in range(steps):
        img, uv = synth.disc_batch(gen, 2, 128, 160, K=4)
        losses.append(train_ops.train_step(m, opt, img.cuda(), uv.cuda(), sigma=8.0, exchange=exchange).item())
    torch.cuda.synchronize()
    return opt.flat_param.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), losses


def main():
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    a = run(True)
    b = run(False)
    ok = all(torch.equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("losses", a[3])
        print("DP step == local step bit for bit on every rank:", bool(flag.item()))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
