"""Import the UNMODIFIED reference (vainaviv/hulk-keypoints) from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the restatement in
`oracle/keypoints_oracle.py` against the real reference classes and (b) generate the golden
vectors under `tests/golden/` (see `oracle/make_golden.py`).  `/root/reference` exists only in
the build container, never on the GPU box, so nothing in `-m gpu` tests, `smoke()` or
`bench.py` may import this file.  The product package never imports anything under `oracle/`.

The reference cannot be imported as-is offline; four shims are needed:
  1. `src/model.py:7-8` does `from resnet_dilated import ...`  -> put `<ref>/src` on sys.path.
  2. `src/resnet_dilated.py:11` hard-codes `pretrained=True` -> `model_zoo.load_url` needs a
     network; we make it return None and make `load_state_dict(None)` a no-op for the duration
     of the constructor so the seeded random init (`src/resnet.py:155-161`) survives.
  3. `src/dataset.py:13` / `src/prediction.py:3` import imgaug / matplotlib (absent, unused).
  4. `src/dataset.py:40,68` call `.cuda()` unconditionally -> identity on a CPU-only host.
"""
import contextlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("HULK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "model.py"))


_installed = False


def _install_shims():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "src"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("imgaug", "imgaug.augmenters", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if isinstance(sys.modules["imgaug"], types.ModuleType) and not hasattr(sys.modules["imgaug"], "augmenters"):
        sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import resnet  # noqa: F401  (<ref>/src/resnet.py)

    resnet.model_zoo.load_url = lambda url, *a, **k: None
    _installed = True


@contextlib.contextmanager
def _tolerate_none_state_dict():
    original = nn.Module.load_state_dict

    def patched(self, sd, *a, **k):
        if sd is None:
            return None
        return original(self, sd, *a, **k)

    nn.Module.load_state_dict = patched
    try:
        yield
    finally:
        nn.Module.load_state_dict = original


def build_reference_model(seed: int, num_keypoints: int = 4, img_height: int = 480, img_width: int = 640):
    """`KeypointsGauss(K, H, W)` of `src/model.py:10-22`, constructed right after `manual_seed(seed)`.

    Returned in train() mode, exactly as the reference constructor leaves it.
    """
    _install_shims()
    from src.model import KeypointsGauss  # reference class

    with _tolerate_none_state_dict():
        torch.manual_seed(seed)
        model = KeypointsGauss(num_keypoints, img_height=img_height, img_width=img_width)
    return model


def reference_modules():
    """Return (dataset_module, prediction_module) of the reference."""
    _install_shims()
    import src.dataset as ds
    import src.prediction as pr

    return ds, pr
