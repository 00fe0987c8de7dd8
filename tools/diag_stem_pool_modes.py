"""Time the fused stem+pool kernel (batch 64, 480x640, fp32 input) with parts switched off (diagnostics build, HK_SP_DEBUG bit flags:
1 = no TMA input loads, 2 = no MMAs, 4 = no epilogue TMEM loads, 8 = converters idle, 16 = epilogue arithmetic / stores off): which role bounds it?  Results of the crippled modes are garbage.
   python -m hulk_keypoints_b200.build --diag ; python tools/diag_stem_pool_modes.py"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:   # child: one mode per process (the flag is read at launch)
    os.environ["HK_LIB_PATH"] = os.path.join(ROOT, "hulk_keypoints_b200", "libhulk_sm100_diag.so")
    sys.path.insert(0, ROOT)
    import torch
    from hulk_keypoints_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.rand(64, 3, 480, 640, generator=g).to(dev)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).to(dev)
    scale, bias = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    wp = ops.stem_pack_weights(w)
    out = torch.empty(64, 120, 160, 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.stem_pool(x, wp, scale, bias, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.stem_pool(x, wp, scale, bias, out=out)
    e1.record(); e1.synchronize()
    print("HK_SP_DEBUG=%s: %.4f ms per launch" % (os.environ.get("HK_SP_DEBUG", "0"), e0.elapsed_time(e1) / 20))
else:
    for mode in ("0", "2", "7", "8", "16", "15", "23", "24", "31"):
        env = dict(os.environ, HK_SP_DEBUG=mode)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
