"""Pick the step count of the k4_480x640 trained fixture: for several totals, train it and report the oracle's argmax margins on exactly
the images tests/test_gpu_parity.py::test_bf16_batch64_480x640_vs_oracle and the train-step tests evaluate.  GPU box.
usage: python tools/diag_ftrn_margin.py [steps ...]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hulk_keypoints_b200 as hk
from hulk_keypoints_b200 import synth
from oracle import keypoints_oracle as O

def margin(heat, kp, excl=7):
    out = []
    for k in range(heat.shape[0]):
        h = heat[k].copy(); y, x = kp[k]; top = h[y, x]
        h[max(0, y - excl): y + excl + 1, max(0, x - excl): x + excl + 1] = -np.inf
        out.append(float(top - h.max()))
    return np.array(out)

name = "k4_480x640"
cfg = synth.FIXTURES[name]
img, uv = synth.disc_batch(torch.Generator().manual_seed(777), 64, 480, 640, 4, on_lattice=True)
idx = [0, 9, 18, 27, 36, 45, 54, 63]
for total in [int(a) for a in sys.argv[1:]] or [400, 450, 500, 550, 600]:
    t0 = time.time()
    torch.manual_seed(cfg["model_seed"])
    model = hk.KeypointsGauss(cfg["K"], img_height=cfg["H"], img_width=cfg["W"]).cuda()
    losses = synth.fit_synthetic(model, total, cfg["B"], cfg["H"], cfg["W"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], sigma=cfg["sigma"],
                                 data_seed=cfg["data_seed"], decay_after=cfg["decay_after"])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ref = O.forward(sd, img[idx], 4).numpy()
    kp_ref = O.argmax_decode(ref)
    mg = np.stack([margin(ref[i], kp_ref[i]) for i in range(len(idx))])
    m = hk.KeypointsGauss(4); m.load_state_dict(sd); m = m.cuda().eval()
    heat, yx = m.heatmaps_and_keypoints(img[idx].cuda())
    d = np.abs(heat.cpu().numpy() - ref)
    print(f"steps {total}: loss {losses[-1]:.5f} margin min {mg.min():.4f} sorted {np.sort(mg.ravel())[:4].round(4).tolist()} peak {ref.reshape(8, 4, -1).max(-1).min():.3f} "
          f"max|d| {d.max():.2e} kp dist {np.abs(yx.cpu().numpy() - kp_ref).max()} loc err {np.abs(kp_ref[..., ::-1] - uv[idx].numpy()).max():.0f} [{time.time() - t0:.1f} s]", flush=True)
