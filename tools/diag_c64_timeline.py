"""Per-role timeline of the c64 conv kernel's CTA 0 (GPU box)."""
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
res = len(sys.argv) > 1 and sys.argv[1] == "res"
x = torch.randn(B, 120, 160, 64, device=dev).to(torch.bfloat16)
w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05)
wp, s, b = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
out = torch.empty_like(x)
r = torch.randn_like(x) if res else None
run = lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=1, dil=1, relu=True, residual=r, out=out)
for _ in range(3): run()
buf = torch.zeros(3 * 16 * 8, device=dev, dtype=torch.int64)
lib.hk_debug_set_c64_timeline.argtypes = [C.c_void_p]
lib.hk_debug_set_c64_timeline(C.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
lib.hk_debug_set_c64_timeline(None)
t = buf.cpu().view(3, 16, 8)
t0 = int(t[1, 0, 0])
rel = lambda v: int(v) - t0
print("PRODUCER per tile: [before wait, after wait] x 3 patches")
for i in range(2, 12): print(i, [rel(v) for v in t[0, i, :6]])
print("MMA per tile: start, after tmem_empty, full s0, s1, s2, issued")
for i in range(2, 12): print(i, [rel(v) for v in t[1, i, :6]])
print("EPILOGUE per tile: start, after bar1, after tmem_full, after res, after drain+arrive, after bar2")
for i in range(2, 12): print(i, [rel(v) for v in t[2, i, :6]])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); e1.synchronize()
print("conv ms:", e0.elapsed_time(e1) / 10)
