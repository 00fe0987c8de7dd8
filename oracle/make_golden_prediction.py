"""Golden vectors for the host-side Prediction helpers, produced by the UNMODIFIED reference class (src/prediction.py:26-38).

TEST INFRASTRUCTURE ONLY (build container; /root/reference must exist):   python -m oracle.make_golden_prediction
Writes tests/golden/prediction_v1.json: seeded (H, W) float32 maps -> Prediction.expectation / softmax checksums.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402


def maps():
    rng = np.random.RandomState(7)
    out = []
    for h, w in ((5, 7), (12, 12), (48, 64), (33, 17)):
        d = rng.rand(h, w).astype(np.float32)
        d[rng.randint(h), rng.randint(w)] += 6.0      # one clear peak, like a trained heatmap
        out.append(d)
    out.append((np.eye(9, 13, dtype=np.float32) * 3).astype(np.float32))
    return out


def main():
    reference_loader._install_shims()
    import importlib
    ref_pred = importlib.import_module("prediction")          # <ref>/src/prediction.py
    p = ref_pred.Prediction(None, 4, 480, 640, False)
    cases = []
    for d in maps():
        sm = p.softmax(d.ravel().astype(np.float64))
        cases.append({"shape": list(d.shape), "expectation": [int(v) for v in p.expectation(d)],
                      "softmax_max": float(sm.max()), "softmax_sum": float(sm.sum())})
    path = os.path.join(ROOT, "tests", "golden", "prediction_v1.json")
    with open(path, "w") as f:
        json.dump({"source": "reference src/prediction.py (unmodified), numpy " + np.__version__, "cases": cases}, f, indent=1)
    print(path, cases)


if __name__ == "__main__":
    main()
