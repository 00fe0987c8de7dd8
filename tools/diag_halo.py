"""A/B of the haloed-operand conv kernel per layer shape (GPU box): HK_CONV_HALO=0 vs 1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import ops
dev = "cuda:0"
def t(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b) / n * 1e3
for name, B, H, W, cin, cout, dil in (("layer2 60x80", 64, 60, 80, 128, 128, 1), ("layer3 60x80", 64, 60, 80, 256, 256, 2), ("layer4 60x80", 64, 60, 80, 512, 512, 4),
                                      ("layer2 120x160", 16, 120, 160, 128, 128, 1), ("layer3 120x160", 16, 120, 160, 256, 256, 2), ("layer4 120x160", 16, 120, 160, 512, 512, 4)):
    x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wp, s, b = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
    out = torch.empty(B, H, W, cout, device=dev, dtype=torch.bfloat16)
    res = torch.randn_like(out)
    outs = {}
    for mode in ("0", "1"):
        os.environ["HK_CONV_HALO"] = mode
        run = lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=dil, dil=dil, relu=True, residual=res, out=out)
        us = t(run)
        outs[mode] = out.clone()
        print(f"{name:16s} halo={mode}: {us:8.1f} us  {2*B*H*W*cout*cin*9/us/1e6:7.1f} TF/s")
    d = (outs["0"].float() - outs["1"].float()).abs().max().item()
    print(f"{name:16s} max |halo - plain| = {d:.4f}")
os.environ.pop("HK_CONV_HALO", None)
