"""How many steps do the trained fixtures need?  Trains each FIXTURES entry with its seeds, and at several step counts evaluates
the bf16 engine against the CPU oracle on lattice discs: oracle peak values, argmax margin (top node vs the best pixel >= 8 px away),
max|d|, keypoint distance.  GPU box.   usage: python tools/diag_ftrn_quality.py [name ...]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hulk_keypoints_b200 as hk
from hulk_keypoints_b200 import synth, train_ops
from hulk_keypoints_b200.optim import FusedAdam
from oracle import keypoints_oracle as O

def margin(heat, kp, excl=7):
    out = []
    for k in range(heat.shape[0]):
        h = heat[k].copy(); y, x = kp[k]; top = h[y, x]
        h[max(0, y - excl): y + excl + 1, max(0, x - excl): x + excl + 1] = -np.inf
        out.append(float(top - h.max()))
    return np.array(out)

PLAN = {"k4_480x640": ([300, 450], (480, 640), [4], 8), "k32_240x320": ([1200, 1800], (960, 1280), [16, 32], 2),
        "k4_128x160": ([300, 450], (128, 160), [4], 4)}
for name in (sys.argv[1:] or list(PLAN)):
    cfg = synth.FIXTURES[name]
    totals, (EH, EW), Ks, EB = PLAN[name]
    for total in totals:
        t0 = time.time()
        torch.manual_seed(cfg["model_seed"])
        model = hk.KeypointsGauss(cfg["K"], img_height=cfg["H"], img_width=cfg["W"]).cuda()
        losses = synth.fit_synthetic(model, total, cfg["B"], cfg["H"], cfg["W"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], sigma=cfg["sigma"],
                                     data_seed=cfg["data_seed"], decay_after=0.6)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        for K in Ks:
            img, uv = synth.disc_batch(torch.Generator().manual_seed(900 + K), EB, EH, EW, K, on_lattice=True)
            ref = O.forward(sd, img, K).numpy()
            m = hk.KeypointsGauss(K); m.load_state_dict(sd); m = m.cuda().eval()
            heat, yx = m.heatmaps_and_keypoints(img.cuda())
            got = heat.cpu().numpy(); kp = yx.cpu().numpy().astype(np.int64); kp_ref = O.argmax_decode(ref)
            mg = np.stack([margin(ref[i], kp_ref[i]) for i in range(EB)])
            peaks = ref.reshape(EB, K, -1).max(-1)
            print(f"{name} total {total} loss {losses[-1]:.5f} eval K={K} {EH}x{EW} B={EB}: peak {peaks.min():.3f}..{peaks.max():.3f} margin min {mg.min():.3f} "
                  f"(n<0.04: {(mg < 0.04).sum()}/{mg.size}) max|d| {np.abs(got - ref).max():.2e} kp dist max {np.abs(kp - kp_ref).max()} "
                  f"loc err {np.abs(kp_ref[..., ::-1] - uv.numpy()).max():.0f}  [{time.time() - t0:.1f} s]", flush=True)
