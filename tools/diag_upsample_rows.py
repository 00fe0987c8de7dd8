"""Per-launch time of the throughput-mode upsample+sigmoid (B=64, K=4, 60x80 -> 480x640) for several rows-per-CTA settings (HK_HEAD_ROWS).
GPU box.  usage: python tools/diag_upsample_rows.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hulk_keypoints_b200 import ops

dev = torch.device("cuda:0")
logits = torch.randn(64, 4, 60, 80, device=dev)
heat = torch.empty(64, 4, 480, 640, device=dev)
ref = None
for rep in range(2):
    for rows in (8, 12, 16, 24, 32, 48, 60, 96):
        os.environ["HK_HEAD_ROWS"] = str(rows)
        for _ in range(3):
            ops.head_upsample(logits, 480, 640, heat=heat)
        torch.cuda.synchronize()
        if ref is None:
            ref = heat.clone()
        assert torch.equal(ref, heat), rows
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.head_upsample(logits, 480, 640, heat=heat)
        e1.record(); e1.synchronize()
        print("rows %3d: %.4f ms  (%.0f GB/s of stores)" % (rows, e0.elapsed_time(e1) / 20, heat.numel() * 4 / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9))
