"""Per-launch time of the last conv (512 -> 512, dilation 4, 60x80, batch 64) as hk_conv_bn_act_fwd and as hk_conv_head_fwd with K = 1, 4, 8
scoring rows: what does the fused epilogue cost?  GPU box.  usage: python tools/diag_conv_head.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hulk_keypoints_b200 import ops

dev = torch.device("cuda:0")
B, H, W, C = 64, 60, 80, 512
g = torch.Generator().manual_seed(1)
x = torch.randn(B, H, W, C, generator=g).clamp_min(0).to(dev).to(torch.bfloat16)
res = torch.randn(B, H, W, C, generator=g).clamp_min(0).to(dev).to(torch.bfloat16)
w = (torch.randn(C, C, 3, 3, generator=g) * (2.0 / (9 * C)) ** 0.5).to(dev)
wp, _, _ = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
s, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
y = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

for rep in range(2):
    print("plain conv  %.4f ms" % timed(lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=4, dil=4, relu=True, residual=res, out=y)))
    for K in (1, 4, 8):
        w_fc, b_fc = torch.randn(K, C, device=dev) * 0.05, torch.zeros(K, device=dev)
        logits = torch.empty(B, K, H, W, device=dev)
        print("fused K=%d   %.4f ms" % (K, timed(lambda: ops.conv_head(x, wp, s, b, w_fc, b_fc, logits, stride=1, pad=4, dil=4, relu=True, residual=res))))
