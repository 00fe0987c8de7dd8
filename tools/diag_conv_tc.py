"""Diagnostics for the tcgen05 conv (GPU box): per-case error summary with structure hints.
Usage: python tools/diag_conv_tc.py [case_index ...]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import _lib, ops  # noqa: E402

CASES = [
    # B, H, W, cin, cout, k, stride, dil
    (1, 4, 16, 64, 64, 1, 1, 1),
    (1, 8, 16, 64, 64, 1, 1, 1),
    (1, 8, 16, 128, 128, 1, 1, 1),
    (1, 8, 16, 64, 256, 1, 1, 1),
    (1, 8, 16, 64, 64, 3, 1, 1),
    (1, 12, 20, 64, 64, 3, 1, 2),
    (2, 30, 40, 64, 128, 3, 2, 1),
    (2, 30, 40, 64, 128, 1, 2, 1),
    (1, 60, 80, 256, 512, 3, 1, 4),
]


def run(case):
    B, H, W, cin, cout, k, stride, dil = case
    g = torch.Generator().manual_seed(1)
    pad = dil * (k - 1) // 2
    x = torch.randn(B, cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, generator=g) * (1.0 / (k * k * cin)) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.double(), w.double(), None, stride, pad, dil)
    dev = torch.device("cuda:0")
    wp, s, b = ops.pack_conv_weights(w.float().to(dev), None, 1e-5, torch.bfloat16)
    y = ops.conv_bn_act(x.permute(0, 2, 3, 1).contiguous().to(dev), wp, s, b, stride=stride, pad=pad, dil=dil, relu=False,
                        algo=_lib.HK_CONV_TCGEN05)
    torch.cuda.synchronize()
    got = y.float().cpu().permute(0, 3, 1, 2).double()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3 * ref.abs().max()
    bad = err > tol
    print(f"case {case}: max|ref|={ref.abs().max():.3f} max err={err.max():.4f} bad={bad.sum().item()}/{bad.numel()}"
          f" got_zero_frac={(got == 0).double().mean():.3f} nan={torch.isnan(got).sum().item()}")
    if bad.any():
        # structure: which channels / rows / cols are wrong
        per_c = bad.sum(dim=(0, 2, 3))
        per_y = bad.sum(dim=(0, 1, 3))
        per_x = bad.sum(dim=(0, 1, 2))
        print("  bad per channel (first 64):", per_c[:64].tolist())
        print("  bad per out row:", per_y.tolist()[:64])
        print("  bad per out col:", per_x.tolist()[:64])
        # try to identify permutations: correlate got with ref under channel-chunk swizzles
        b0 = 0
        gv, rv = got[b0, :, 0, 0], ref[b0, :, 0, 0]
        print("  got[0,:8,0,0]", [round(v, 3) for v in gv[:8].tolist()])
        print("  ref[0,:8,0,0]", [round(v, 3) for v in rv[:8].tolist()])
        # does some other pixel's reference match this pixel's output? (M-row permutation)
        flat_ref = ref[b0].reshape(ref.shape[1], -1)
        for pix in (0, 1, 8, 16, 17):
            if pix < flat_ref.shape[1]:
                gp = got[b0].reshape(got.shape[1], -1)[:, pix]
                d = (flat_ref - gp[:, None]).abs().mean(dim=0)
                j = int(d.argmin())
                print(f"  out pixel {pix} best matches ref pixel {j} (mean abs diff {d[j]:.4f}; own {d[pix]:.4f})")
    return not bad.any()


if __name__ == "__main__":
    _lib.require_device()
    idx = [int(a) for a in sys.argv[1:]] or range(len(CASES))
    ok = True
    for i in idx:
        try:
            ok &= run(CASES[i])
        except Exception as e:  # keep going: later cases may still tell something
            print(f"case {CASES[i]}: EXCEPTION {e}")
            ok = False
            break
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
