"""Which side bounds the 2-CTA conv mainloop?  Times each layer shape with HK_TC2_DEBUG = 0 (normal), 1 (no MMA), 2 (no TMA)."""
import os, sys
import torch
# needs the diagnostics build of the library (timeline stamps / HK_TC2_DEBUG switches are compiled out of the shipped .so):
#   python -m hulk_keypoints_b200.build --diag   ->  hulk_keypoints_b200/libhulk_sm100_diag.so
os.environ.setdefault("HK_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hulk_keypoints_b200", "libhulk_sm100_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import ops
dev = torch.device("cuda:0")
B = 64
for name, cin, cout, dil, res in (("layer2.c1", 128, 128, 1, False), ("layer3.c1", 256, 256, 2, False), ("layer4.c1", 512, 512, 4, False)):
    x = torch.randn(B, 60, 80, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    wp, s, b = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
    out = torch.empty(B, 60, 80, cout, device=dev, dtype=torch.bfloat16)
    run = lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=dil, dil=dil, relu=True, residual=None, out=out)
    for mode in ("0", "1", "2", "4", "6", "5"):
        os.environ["HK_TC2_DEBUG"] = mode
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} mode={mode}: {ms*1e3:.1f} us  ({2*B*4800*cout*cin*9/ms/1e9:.1f} TF/s equivalent)")
os.environ["HK_TC2_DEBUG"] = "0"
