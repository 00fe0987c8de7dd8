"""Checkpoint interop.

* `save_checkpoint` / `load_checkpoint`: the reference's format -- a plain `state_dict` pickle with its 218 keys
  (train.py:47-48, analysis.py:19) -- so files interchange with the reference in both directions.
* `load_pretrained_backbone`: the reference initialises the backbone from torchvision's `resnet34-333f7ec4.pth`
  (`resnet34(pretrained=True)`, src/resnet.py:13,237-238) and then replaces `fc` (src/resnet_dilated.py:16).  There is no
  network here, so the same file (or any torchvision-format ResNet-34 state dict) can be supplied offline; `fc.*` is
  ignored exactly as the reference discards it.
"""
from __future__ import annotations

from typing import Dict, Union

import torch

PREFIX = "resnet.resnet34_8s."


def save_checkpoint(model: torch.nn.Module, path: str) -> None:
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)


def load_checkpoint(model: torch.nn.Module, path: str) -> None:
    sd = torch.load(path, map_location="cpu")
    model.load_state_dict(sd)


def find_pretrained_resnet34():
    """Path of `resnet34-333f7ec4.pth` (the file reference src/resnet.py:13 downloads) in the torch hub cache, or None."""
    import os

    name = "resnet34-333f7ec4.pth"
    roots = [os.path.join(torch.hub.get_dir(), "checkpoints"), os.path.expanduser("~/.torch/models")]
    for r in roots:
        p = os.path.join(r, name)
        if os.path.exists(p):
            return p
    return None


def load_pretrained_backbone(model: torch.nn.Module, source: Union[str, Dict[str, torch.Tensor]]) -> int:
    """Copy a torchvision-format ResNet-34 state dict (keys `conv1.weight`, `layer3.0.downsample.1.running_var`, ...)
    into the backbone.  Returns the number of tensors loaded.  `fc.*` is skipped (the reference replaces it,
    src/resnet_dilated.py:16).  `*.num_batches_tracked` may be absent -- `resnet34-333f7ec4.pth` predates that buffer -- exactly as
    torch's own BatchNorm loader tolerates it.  Keys and shapes are validated BEFORE anything is copied: a bad file leaves the
    model untouched."""
    sd = torch.load(source, map_location="cpu") if isinstance(source, str) else source
    own = model.state_dict()
    plan = []
    for k, v in sd.items():
        if k.startswith("fc."):
            continue
        tk = PREFIX + k
        if tk not in own:
            raise KeyError(f"unexpected key {k!r} in the pretrained state dict")
        if tuple(own[tk].shape) != tuple(v.shape):
            raise ValueError(f"shape mismatch for {k}: {tuple(v.shape)} vs {tuple(own[tk].shape)}")
        plan.append((own[tk], v))
    missing = [k for k in own if not k.startswith(PREFIX + "fc.") and not k.endswith("num_batches_tracked")
               and k[len(PREFIX):] not in sd]
    if missing:
        raise KeyError(f"pretrained state dict lacks {len(missing)} backbone tensors, e.g. {missing[:3]}")
    with torch.no_grad():
        for dst, v in plan:
            dst.copy_(v)
    if hasattr(model, "mark_weights_changed"):
        model.mark_weights_changed()
    return len(plan)
