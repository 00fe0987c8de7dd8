"""Stand-alone timing of the train step's BatchNorm accumulator kernels (hk_bn_stats_acc, hk_bn_apply_fwd_acc, hk_bn_bwd_acc) at the
shapes of a per-GPU batch-32 (or --batch N) 480x640 train step, one library build per run (HK_LIB_PATH selects it).

    HK_LIB_PATH=$PWD/tools/bin/libhulk_bnA.so python tools/diag_bn_kernels.py [--batch 32] [--iters 20] [--json out.json]

Every call works on the next of several buffer sets (together larger than the 126 MB L2), so the numbers are HBM numbers.  GB/s =
algorithmic bytes (every operand read or written once) / mean call duration (CUDA events around `iters` back-to-back calls).  Under ncu
(--iters 1) the launch list gives the per-kernel split of the two-launch backward call.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--shapes", default="stem,layer1,layer2,layer3,layer4")
    args = ap.parse_args()
    B = args.batch
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    # (name, pixels per image, channels): the stem map and the four stages of ResnetDilated-34 at 480x640
    shapes = [("stem", 240 * 320, 64), ("layer1", 120 * 160, 64), ("layer2", 60 * 80, 128), ("layer3", 60 * 80, 256), ("layer4", 60 * 80, 512)]
    shapes = [sh for sh in shapes if sh[0] in args.shapes.split(",")]
    res = {"lib": os.environ.get("HK_LIB_PATH", "default"), "batch": B, "rows": []}
    for name, ppi, C in shapes:
        P = B * ppi
        n = P * C
        sets = max(2, int(200e6 // (n * 2 * 4)) + 1)   # >= ~200 MB of distinct operands per round of calls
        y = [torch.randn(P, C, device=dev).mul_(0.7).add_(0.3).to(torch.bfloat16) for _ in range(sets)]
        d = [torch.randn(P, C, device=dev).mul_(0.01).to(torch.bfloat16) for _ in range(sets)]
        rsd = [torch.randn(P, C, device=dev).to(torch.bfloat16) for _ in range(sets)]
        out = [torch.empty(P, C, device=dev, dtype=torch.bfloat16) for _ in range(sets)]
        dm = [torch.empty(P, C, device=dev, dtype=torch.bfloat16) for _ in range(sets)]
        bits = [torch.randint(0, 256, (n // 8,), device=dev, dtype=torch.uint8) for _ in range(sets)]
        acc = torch.zeros(ops.bn_acc_bytes(C), device=dev, dtype=torch.uint8)
        gamma = torch.rand(C, device=dev) + 0.5
        beta = torch.randn(C, device=dev) * 0.1
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        mean, invstd = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)

        def timed(fn, nbytes):
            for k in range(2):
                fn(k % sets)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(args.iters):
                fn(k % sets)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            return ms * 1e3, nbytes / ms / 1e6

        def f_stats(k):
            acc.zero_()
            ops.bn_stats_acc(y[k], acc)

        def f_apply(k, residual):
            ops.bn_apply_acc(y[k], acc, gamma, beta, rm, rv, 0.1, 1e-5, mean, invstd, True, residual=rsd[k] if residual else None, out=out[k], relu_bits=bits[k])

        def f_bwd(k, dmasked):
            acc.zero_()
            ops.bn_bwd_acc(d[k], bits[k], y[k], mean, invstd, gamma, acc, dg, db, out[k], dmasked=dm[k] if dmasked else None)

        acc.zero_()
        ops.bn_stats_acc(y[0], acc)   # real statistics for the apply calls
        rows = [
            ("stats", timed(f_stats, n * 2)),
            ("apply_fwd", timed(lambda k: f_apply(k, False), n * 4 + n // 8)),
            ("apply_fwd+res", timed(lambda k: f_apply(k, True), n * 6 + n // 8)),
            ("bwd(reduce+apply)", timed(lambda k: f_bwd(k, False), n * 10 + n // 4)),
            ("bwd+dmasked", timed(lambda k: f_bwd(k, True), n * 12 + n // 4)),
        ]
        for what, (us, gbs) in rows:
            print(f"{name:7s} P={P:8d} C={C:4d} {what:18s} {us:8.1f} us {gbs:7.0f} GB/s")
            res["rows"].append({"shape": name, "P": P, "C": C, "what": what, "us": us, "gbs": gbs})
        del y, d, rsd, out, dm, bits
        torch.cuda.empty_cache()
    tot = sum(r["us"] for r in res["rows"])
    print(f"sum of rows: {tot:.1f} us  ({res['lib']})")
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
