"""Phase timeline of the stem kernel's CTA 0 (GPU box).  Prints per-tile cycle deltas between phase stamps."""
import ctypes as C, os, sys
import torch
# needs the diagnostics build of the library (timeline stamps / HK_TC2_DEBUG switches are compiled out of the shipped .so):
#   python -m hulk_keypoints_b200.build --diag   ->  hulk_keypoints_b200/libhulk_sm100_diag.so
os.environ.setdefault("HK_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hulk_keypoints_b200", "libhulk_sm100_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import _lib, ops
lib = _lib.lib()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand(B, 3, 480, 640, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
s, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
wp = ops.stem_pack_weights(w)
out = torch.empty((B, 240, 320, 64), device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.stem(x, wp, s, b, out=out)
buf = torch.zeros(24 * 8 + 8, device=dev, dtype=torch.int64)
lib.hk_debug_set_stem_timeline.argtypes = [C.c_void_p]
lib.hk_debug_set_stem_timeline(C.c_void_p(buf.data_ptr()))
ops.stem(x, wp, s, b, out=out)
torch.cuda.synchronize()
lib.hk_debug_set_stem_timeline(None)
whole = buf.cpu()
t = whole[:192].view(24, 8)
for name, o in (("CTA 0", 192), ("last CTA", 196)):
    print(f"{name}: loop {int(whole[o+1]-whole[o])} clk, final store drain {int(whole[o+2]-whole[o+1])} clk, tiles {int(whole[o+3])}")
print('CTA0 entry -> last CTA exit', int(whole[198]-whole[192]))
names = ["sts_patch", "sync1", "build", "fence+sync2", "mma_issue+prefetch", "mma_wait", "epilogue", "loop"]
print("tile  " + " ".join(f"{n:>18s}" for n in names))
for i in range(2, 20):
    d = [int(t[i, j + 1] - t[i, j]) for j in range(7)] + [int(t[i + 1, 0] - t[i, 7])]
    print(f"{i:4d}  " + " ".join(f"{v:18d}" for v in d) + f"   total {int(t[i+1,0]-t[i,0])}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.hk_debug_set_stem_timeline(C.c_void_p(buf.data_ptr()))
for _ in range(10):
    ops.stem(x, wp, s, b, out=out)
e1.record(); e1.synchronize()
print("stem ms:", e0.elapsed_time(e1) / 10)
