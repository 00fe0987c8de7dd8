"""Per-role timeline of the 2-CTA conv kernel's cluster-0 leader (GPU box).  usage: diag_tc2_timeline.py"""
import ctypes as C, os, sys
import torch
# needs the diagnostics build of the library (timeline stamps / HK_TC2_DEBUG switches are compiled out of the shipped .so):
#   python -m hulk_keypoints_b200.build --diag   ->  hulk_keypoints_b200/libhulk_sm100_diag.so
os.environ.setdefault("HK_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hulk_keypoints_b200", "libhulk_sm100_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import _lib, ops
lib = _lib.lib()
dev = torch.device("cuda:0")
B = 64
lib.hk_debug_set_tc2_timeline.argtypes = [C.c_void_p]
SHAPES = (("layer2.c1", 128, 128, 1, False),) if os.environ.get("HK_TC2_DEBUG") else (("layer2.c1", 128, 128, 1, False), ("layer2.c2", 128, 128, 1, True), ("layer3.c1", 256, 256, 2, False),
                                  ("layer3.c2", 256, 256, 2, True), ("layer3.0.c1", 128, 256, 2, False), ("layer4.c2", 512, 512, 4, True))
for name, cin, cout, dil, res in SHAPES:
    x = torch.randn(B, 60, 80, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    wp, s, b = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
    out = torch.empty(B, 60, 80, cout, device=dev, dtype=torch.bfloat16)
    r = torch.randn_like(out) if res else None
    run = lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=dil, dil=dil, relu=True, residual=r, out=out)
    for _ in range(3): run()
    buf = torch.zeros(3 * 16 * 8, device=dev, dtype=torch.int64)
    lib.hk_debug_set_tc2_timeline(C.c_void_p(buf.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.hk_debug_set_tc2_timeline(None)
    t = buf.cpu().view(3, 16, 8)
    t0 = int(t[1, 0, 0])
    rel = lambda v: int(v) - t0
    print(f"==== {name}: cin={cin} cout={cout} res={res}")
    print("PRODUCER per tile: start, all issued")
    for i in range(0, 10): print(i, [rel(v) for v in t[0, i, :2]])
    print("MMA per tile: start, after tmem_empty, first full, all issued")
    for i in range(0, 10): print(i, [rel(v) for v in t[1, i, :4]])
    print("EPILOGUE per tile: start, after tmem_full, c0 after bar1, c0 after tmem ld, c0 after res wait, c0 after math, c0 after bar2, last chunk after bar2")
    for i in range(0, 10): print(i, [rel(v) for v in t[2, i, :8]])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"conv ms: {ms:.4f}  TF/s: {2*B*4800*cout*cin*9/ms/1e9:.1f}")
