import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    arrays = dict(np.load(os.path.join(d, "golden_v1.npz")))
    with open(os.path.join(d, "golden_v1.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def golden2():
    """Round-2 goldens from the unmodified reference (oracle/make_golden_v2.py): KeypointsDataset round trip, normalize_dist,
    Prediction.expectation before truncation."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz")))


def sd_digest(sd) -> str:
    import hashlib

    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rand_img(seed, b, h, w):
    import torch

    return torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(seed))
