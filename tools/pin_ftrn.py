"""Train the deterministic F-trn fixtures of tests/test_gpu_parity.py on the GPU box and write their SHA-256 pins.

    gpurun -- 'python tools/pin_ftrn.py'      ->  gpurun_out/ftrn_v2.json   (copy it to tests/golden/ftrn_v2.json and commit)

Re-run after any change to a kernel of the training step (the fixture bytes follow the kernels' rounding)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hulk_keypoints_b200 import synth

out = {"sha256": {}, "losses": {}, "how": "tools/pin_ftrn.py: hulk_keypoints_b200.synth.train_fixture(name) on one B200", "fixtures": synth.FIXTURES}
for name in synth.FIXTURES:
    t0 = time.time()
    sd, losses = synth.train_fixture(name)
    out["sha256"][name] = synth.state_dict_sha256(sd)
    out["losses"][name] = {"first": losses[0], "last": losses[-1], "min": min(losses)}
    print(name, out["sha256"][name], f"loss {losses[0]:.4f} -> {losses[-1]:.5f}", f"{time.time() - t0:.1f} s", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for p in (os.path.join(ROOT, "gpurun_out", "ftrn_v2.json"), os.path.join(ROOT, "tests", "golden", "ftrn_v2.json")):
    with open(p, "w") as f:
        json.dump(out, f, indent=1)
print("wrote pins")
