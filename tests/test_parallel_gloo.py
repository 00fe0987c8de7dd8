"""CPU, world_size 2, gloo: the N>1 host logic (shards, gradient exchange, keypoint gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hulk_keypoints_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 2, 1))
        if rank == 1:  # de-synchronise, then broadcast must repair
            for p in model.parameters():
                p.data.add_(1.0)
        parallel.broadcast_model(model, src=0)
        ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 2, 1))
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 2, 1))
        same_after_bcast = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))

        # data-parallel step on a global batch of 6 split 3/3: averaged shard grads == full-batch grads
        # (BN in eval so per-replica statistics do not enter)
        model.eval()
        ref.eval()
        g = torch.Generator().manual_seed(5)
        x = torch.rand(6, 3, 10, 10, generator=g)
        b, e = parallel.shard_range(6, world, rank)
        model(x[b:e]).mean().backward()
        calls = parallel.allreduce_gradients(model.parameters(), bucket_bytes=256)
        ref(x).mean().backward()
        grads_ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(model.parameters(), ref.parameters()))

        # keypoint gather keeps shard order, uneven shards
        total = 5
        b, e = parallel.shard_range(total, world, rank)
        local = torch.arange(b, e, dtype=torch.int32).view(-1, 1, 1).repeat(1, 4, 2)
        full = parallel.gather_keypoints(local, total)
        gather_ok = full.shape == (total, 4, 2) and torch.equal(full[:, 0, 0], torch.arange(total, dtype=torch.int32))
        result_q.put((rank, same_after_bcast, grads_ok, calls, gather_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, same, grads_ok, calls, gather_ok in results:
        assert same, f"rank {rank}: broadcast_model did not synchronise"
        assert grads_ok, f"rank {rank}: averaged gradients differ from the full-batch gradients"
        assert calls >= 2, "small bucket size must produce several collectives"
        assert gather_ok, f"rank {rank}: gather_keypoints order"


def test_single_process_is_noop():
    m = torch.nn.Linear(2, 2)
    m(torch.ones(1, 2)).sum().backward()
    assert parallel.allreduce_gradients(m.parameters()) == 0
    t = torch.zeros(3, 4, 2, dtype=torch.int32)
    assert parallel.gather_keypoints(t, 3) is t
