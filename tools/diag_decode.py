"""Stand-alone timing of hk_argmax_decode (HK_LIB_PATH selects the build): GB/s = heatmap bytes / mean call time, CUDA events around
`iters` calls rotating over two heatmap buffers (each larger than the 126 MB L2).
    python tools/diag_decode.py [--iters 20]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
for (B, K, H, W) in [(64, 4, 480, 640), (16, 32, 960, 1280), (4, 4, 480, 640)]:
    heat = [torch.rand(B, K, H, W, device=dev) for _ in range(2)]
    yx = torch.empty(B, K, 2, device=dev, dtype=torch.int32)
    mv = torch.empty(B, K, device=dev)
    ws = torch.empty(ops.argmax_workspace_bytes(B, K, H, W), device=dev, dtype=torch.uint8)
    ref = [h.view(B * K, -1).argmax(1) for h in heat]
    for k in range(2):
        ops.argmax_decode(heat[k], yx, mv, ws)
        got = yx.view(-1, 2)[:, 0].long() * W + yx.view(-1, 2)[:, 1].long()
        assert torch.equal(got, ref[k]), "argmax mismatch"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.iters):
        ops.argmax_decode(heat[k & 1], yx, mv, ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"decode B={B} K={K} {H}x{W}: {ms * 1e3:7.1f} us  {heat[0].numel() * 4 / ms / 1e6:6.0f} GB/s  ({os.environ.get('HK_LIB_PATH', 'default')})")

# Gaussian heatmap targets (reference src/dataset.py:36-44): fp64 (B,K,H,W) written once
for (B, K, H, W) in [(64, 4, 480, 640), (4, 4, 480, 640)]:
    uv = torch.rand(B, K, 2, device=dev) * torch.tensor([W, H], device=dev)
    out = torch.empty(B, K, H, W, device=dev, dtype=torch.float64)
    for _ in range(2):
        ops.gauss_targets(uv, H, W, 8.0, torch.float64, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        ops.gauss_targets(uv, H, W, 8.0, torch.float64, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"gauss_targets B={B} K={K} {H}x{W} fp64: {ms * 1e3:7.1f} us  {out.numel() * 8 / ms / 1e6:6.0f} GB/s")
