"""CPU: host-side mirror of the reference API, and the C-ABI library's symbol table (no kernels run)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
import torch

from conftest import ROOT, rand_img, sd_digest
import hulk_keypoints_b200 as hk
from hulk_keypoints_b200 import _lib, parallel
from oracle import keypoints_oracle as O

warnings.filterwarnings("ignore")


def test_constructor_attributes_and_state_dict_keys(golden):
    _, meta = golden
    torch.manual_seed(0)
    m = hk.KeypointsGauss(4, img_height=480, img_width=640)
    assert (m.num_keypoints, m.num_outputs, m.img_height, m.img_width) == (4, 4, 480, 640)
    sd = m.state_dict()
    ref = O.init_state_dict(0)
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 218
    assert all(sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype for k in ref)
    assert sum(p.numel() for p in m.parameters()) == 21_797_672
    assert sd_digest(sd) == sd_digest(ref)  # same seeded init as the oracle (hence as the reference)
    if sd_digest(ref) == meta["weights_sha256_seed0"]:
        assert sd_digest(sd) == meta["weights_sha256_seed0"]


def test_construction_consumes_rng_like_reference():
    torch.manual_seed(7)
    hk.KeypointsGauss(4)
    after_model = torch.rand(1).item()
    torch.manual_seed(7)
    O.init_state_dict(7)
    assert torch.rand(1).item() == after_model


def test_load_state_dict_roundtrip():
    m = hk.KeypointsGauss(4)
    sd = O.init_state_dict(1)
    missing = m.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert sd_digest(m.state_dict()) == sd_digest(sd)


def test_train_mode_forward_matches_oracle_on_cpu():
    torch.manual_seed(0)
    m = hk.KeypointsGauss(4)
    m.train()
    m.train_backend = "autograd"   # the explicit torch-autograd checker graph (the engine is CUDA-only; nothing falls back silently)
    x = rand_img(6, 2, 64, 96)
    sd = O.init_state_dict(0)
    with torch.no_grad():
        y = m(x)
    y_ref = O.forward(sd, x, 4, train=True)
    assert y.shape == (2, 4, 64, 96)
    assert (y - y_ref).abs().max().item() < 1e-6
    # running stats advanced exactly like the reference's train-mode BN
    stats = {}
    O.forward(sd, x, 4, train=True, new_stats=stats)
    got = m.state_dict()
    assert all(torch.allclose(got[k], v, atol=1e-6) for k, v in stats.items())


def test_dead_fc_rows_get_zero_grad():
    m = hk.KeypointsGauss(4)
    m.train()
    m.train_backend = "autograd"
    m(rand_img(1, 1, 32, 32)).sum().backward()
    g = m.resnet.resnet34_8s.fc.weight.grad
    assert g[:4].abs().sum() > 0 and g[4:].abs().sum() == 0


def test_eval_forward_refuses_cpu():
    m = hk.KeypointsGauss(4).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(rand_img(1, 1, 32, 32))


def test_train_forward_never_falls_back_silently():
    """train() mode is engine-backed; a CPU tensor, the fp32 mode or an odd shape raise instead of quietly running torch/cuDNN."""
    m = hk.KeypointsGauss(4).train()
    with pytest.raises(RuntimeError, match="silent"):
        m(rand_img(1, 1, 32, 32))
    assert "CPU tensor" in m._train_engine_unsupported(torch.zeros(1, 3, 32, 32))
    m.train_backend = "cudnn"
    with pytest.raises(ValueError):
        m(rand_img(1, 1, 32, 32))


def test_bad_arguments():
    with pytest.raises(ValueError):
        hk.KeypointsGauss(0)
    with pytest.raises(ValueError):
        hk.KeypointsGauss(4, precision="fp16")


def test_prediction_shapes_and_eval_switch():
    m = hk.KeypointsGauss(4)
    assert m.training
    calls = []
    m.forward = lambda x: calls.append(tuple(x.shape)) or x
    p = hk.Prediction(m, 4, 480, 640, use_cuda=False)
    assert not m.training  # documented deviation: inference folds BN
    p.predict(torch.zeros(3, 8, 8))
    p.predict(torch.zeros(2, 3, 8, 8))
    assert calls == [(1, 3, 8, 8), (2, 3, 8, 8)]
    with pytest.raises(ValueError):
        p.predict(torch.zeros(8, 8))
    assert p.expectation(np.eye(5, 7, dtype=np.float32) * 50)[0] in range(7)
    m2 = hk.KeypointsGauss(4)
    hk.Prediction(m2, 4, 480, 640, False, bn_mode="as_written")
    assert m2.training


def test_transform_matches_totensor():
    img = (np.random.RandomState(0).rand(6, 5, 3) * 255).astype(np.uint8)
    t = hk.transform(img)
    assert t.shape == (3, 6, 5) and t.dtype == torch.float32
    assert torch.equal(t, torch.from_numpy(img).permute(2, 0, 1).float().div(255))


def test_u8_scale_by_reciprocal_is_exact_in_bf16():
    """The tensor-core stems scale uint8 pixels by the fp32 constant 1/255 instead of dividing (csrc/hk_common.cuh kInv255): after the bf16
    operand rounding the two agree for every byte value, so the uint8 path stays bit-identical to ToTensor (reference src/dataset.py:16)
    followed by the bf16 rounding."""
    b = np.arange(256, dtype=np.float32)
    div = torch.from_numpy(b / np.float32(255.0)).to(torch.bfloat16)
    mul = torch.from_numpy(b * (np.float32(1.0) / np.float32(255.0))).to(torch.bfloat16)
    assert torch.equal(div, mul)


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hulk_sm100.h")).read()
    declared = re.findall(r"HK_API\s+[\w\s\*]+?\b(hk_\w+)\s*\(", header)
    assert len(declared) >= 12 and set(declared) == set(_lib.EXPORTS)
    if not os.path.exists(_lib.LIB_PATH):
        from hulk_keypoints_b200.build import build
        build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported"
    assert _lib.lib().hk_version() == _lib.ABI_VERSION


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhulk_sm100.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    from hulk_keypoints_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.argmax_decode(torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gauss_targets(torch.zeros(1, 1, 2), 4, 4, 1.0)


def test_dropin_shims_reexport_the_package(monkeypatch):
    """`from src.model import KeypointsGauss` etc. (train.py:11-12, analysis.py:8-10) resolve to our classes when the
    dropin/ directory stands in for the reference's src/ (INTEGRATION.md §1)."""
    import importlib
    import sys
    monkeypatch.syspath_prepend(os.path.join(ROOT, "dropin"))
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.") or n == "config"]:
        monkeypatch.delitem(sys.modules, name)
    model = importlib.import_module("src.model")
    dataset = importlib.import_module("src.dataset")
    prediction = importlib.import_module("src.prediction")
    config = importlib.import_module("config")
    assert model.KeypointsGauss is hk.KeypointsGauss
    assert dataset.KeypointsDataset is hk.KeypointsDataset and dataset.transform is hk.transform
    assert dataset.gauss_2d_batch is hk.gauss_2d_batch and prediction.Prediction is hk.Prediction
    assert (config.NUM_KEYPOINTS, config.IMG_HEIGHT, config.IMG_WIDTH, config.GAUSS_SIGMA) == (4, 480, 640, 8)
    m = model.KeypointsGauss(config.NUM_KEYPOINTS, img_height=config.IMG_HEIGHT, img_width=config.IMG_WIDTH)
    assert len(m.state_dict()) == 218


def test_checkpoint_roundtrip_and_pretrained_import(tmp_path):
    from hulk_keypoints_b200 import checkpoint
    torch.manual_seed(3)
    m = hk.KeypointsGauss(4)
    path = str(tmp_path / "model_2_1_0.pth")
    checkpoint.save_checkpoint(m, path)                       # train.py:47-48 format
    sd = torch.load(path)
    assert list(sd.keys()) == list(O.init_state_dict(0).keys())
    m2 = hk.KeypointsGauss(4)
    checkpoint.load_checkpoint(m2, path)                      # analysis.py:19
    assert sd_digest(m2.state_dict()) == sd_digest(m.state_dict())
    # torchvision-format ResNet-34 weights (what resnet34(pretrained=True) downloads, resnet.py:237-238)
    import torchvision
    tv = torchvision.models.resnet34(weights=None)
    n = checkpoint.load_pretrained_backbone(m2, tv.state_dict())
    assert n == len(tv.state_dict()) - 2                      # everything but fc.weight / fc.bias
    own = m2.state_dict()
    assert torch.equal(own[checkpoint.PREFIX + "layer3.0.downsample.0.weight"], tv.state_dict()["layer3.0.downsample.0.weight"])
    assert torch.equal(own[checkpoint.PREFIX + "fc.weight"], m.state_dict()[checkpoint.PREFIX + "fc.weight"])  # untouched
    bad = dict(tv.state_dict())
    bad.pop("bn1.weight")
    with pytest.raises(KeyError):
        checkpoint.load_pretrained_backbone(m2, bad)
    # the file the reference downloads (resnet34-333f7ec4.pth) predates `num_batches_tracked`: an old-format dict loads, and a bad
    # dict leaves the model untouched (validation happens before the first copy)
    old_fmt = {k: v for k, v in tv.state_dict().items() if not k.endswith("num_batches_tracked")}
    m3 = hk.KeypointsGauss(4)
    assert checkpoint.load_pretrained_backbone(m3, old_fmt) == len(old_fmt) - 2
    assert torch.equal(m3.state_dict()[checkpoint.PREFIX + "conv1.weight"], tv.state_dict()["conv1.weight"])
    m4 = hk.KeypointsGauss(4)
    before = sd_digest(m4.state_dict())
    wrong_shape = dict(old_fmt)
    wrong_shape["layer4.2.conv2.weight"] = torch.zeros(512, 512, 1, 1)
    with pytest.raises(ValueError):
        checkpoint.load_pretrained_backbone(m4, wrong_shape)
    assert sd_digest(m4.state_dict()) == before
    # the drop-in constructor reproduces the reference's real initialisation when given the file
    ppath = str(tmp_path / "resnet34-333f7ec4.pth")
    torch.save(old_fmt, ppath)
    m5 = hk.KeypointsGauss(4, pretrained=ppath)
    assert torch.equal(m5.state_dict()[checkpoint.PREFIX + "layer1.0.conv1.weight"], tv.state_dict()["layer1.0.conv1.weight"])


def test_conv_flops_per_image_matches_oracle_and_survey():
    """The FLOP count bench.py's roofline uses (product package) == the oracle's network spec == SURVEY §8d's 211.91 GF."""
    from hulk_keypoints_b200.engine import conv_flops_per_image
    from oracle import keypoints_oracle as O
    for (h, w, k) in ((480, 640, 4), (960, 1280, 16), (64, 96, 7)):
        assert conv_flops_per_image(hk.KeypointsGauss(k, img_height=h, img_width=w), h, w) == O.conv_flops_per_image(h, w, k)
    assert abs(conv_flops_per_image(hk.KeypointsGauss(4), 480, 640) / 1e9 - 211.91) < 0.01
