"""Round-2 golden vectors, produced by the UNMODIFIED reference classes (build container only; /root/reference must exist):

    python -m oracle.make_golden_v2        ->  tests/golden/golden_v2.npz

TEST INFRASTRUCTURE ONLY.  Contents:
  * `KeypointsDataset` (reference src/dataset.py:52-79) over a temporary folder of two 48x64 JPEGs + `%05d.npy` labels, some of them
    outside the image: the JPEG bytes, the raw labels, the clipped labels the reference stores (x clipped to [0, W-1], y to [0, H-1],
    (x, y) order), the image tensor `transform(cv2.imread(path))` and the float64 Gaussians `__getitem__` returns (sigma = 3).
  * `gauss_2d_batch(..., normalize_dist=True)` (src/dataset.py:33-34,42-44).
  * one training step as written (train.py:18-26,35) on seed-0 weights, a (2,3,64,96) batch: float64 loss and a sample of the
    parameter gradients (autograd through the unmodified reference modules).
  * `Prediction.expectation` BEFORE the int() truncation on the maps of make_golden_prediction.maps() plus a 48x64 peaked map, so the
    hk_soft_argmax kernel can be compared in floating point.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402
from oracle.make_golden_prediction import maps  # noqa: E402

H, W, K, SIGMA = 48, 64, 4, 3


def raw_labels():
    # (x, y) pairs; several outside the image on purpose (the reference clips, dataset.py:65-66)
    return [np.array([[10.0, 5.0], [70.0, 20.0], [-3.0, 47.0], [31.5, 60.0]]),
            np.array([[0.0, 0.0], [63.0, 47.0], [63.9, -1.0], [12.25, 30.75]])]


def jpeg_bytes():
    import cv2
    rng = np.random.RandomState(11)
    out = []
    for _ in range(2):
        img = (rng.rand(H, W, 3) * 255).astype(np.uint8)
        img = cv2.GaussianBlur(img, (5, 5), 0)       # smooth content so the JPEG is small
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        assert ok
        out.append(np.frombuffer(buf.tobytes(), dtype=np.uint8))
    return out


def main():
    ds, pr = reference_loader.reference_modules()
    out = {}
    labels, jpgs = raw_labels(), jpeg_bytes()
    with tempfile.TemporaryDirectory() as tmp:
        img_dir, lab_dir = os.path.join(tmp, "images"), os.path.join(tmp, "keypoints")
        os.makedirs(img_dir); os.makedirs(lab_dir)
        for i, (lab, jb) in enumerate(zip(labels, jpgs)):
            np.save(os.path.join(lab_dir, "%05d.npy" % i), lab)
            with open(os.path.join(img_dir, "%05d.jpg" % i), "wb") as f:
                f.write(jb.tobytes())
        dataset = ds.KeypointsDataset(img_dir, lab_dir, K, H, W, ds.transform, gauss_sigma=SIGMA)
        assert len(dataset) == 2
        for i in range(2):
            img, gauss = dataset[i]
            out[f"ds_jpeg_{i}"] = jpgs[i]
            out[f"ds_raw_label_{i}"] = labels[i]
            out[f"ds_clipped_label_{i}"] = dataset.labels[i].cpu().numpy()
            out[f"ds_img_{i}"] = img.numpy()
            out[f"ds_gauss_{i}"] = gauss.cpu().numpy()
            assert img.dtype == torch.float32 and tuple(img.shape) == (3, H, W)
            assert gauss.dtype == torch.float64 and tuple(gauss.shape) == (K, H, W)
    U = torch.tensor([10.0, 63.0, 0.0, 31.5])
    V = torch.tensor([5.0, 20.0, 47.0, 12.25])
    out["gauss_norm_uv"] = torch.stack([U.clone(), V.clone()], -1).numpy()
    out["gauss_norm"] = ds.gauss_2d_batch(W, H, SIGMA, U.clone(), V.clone(), normalize_dist=True).cpu().numpy()

    # Prediction.expectation before truncation: the reference's own arithmetic (prediction.py:26-38) with the final int() removed
    p = pr.Prediction(None, K, H, W, False)
    ms = maps()
    rng = np.random.RandomState(23)
    peaked = (rng.rand(48, 64) * 0.05).astype(np.float32)
    peaked[30, 41] = 0.93
    peaked[29:32, 40:43] += 0.3
    ms.append(peaked)
    for i, d in enumerate(ms):
        width, height = d.T.shape
        flat = d.T.ravel()
        d_norm = p.softmax(flat)
        x_indices = np.array([j % width for j in range(width * height)])
        y_indices = np.array([j // width for j in range(width * height)])
        raw = np.array([np.dot(d_norm, x_indices), np.dot(d_norm, y_indices)], dtype=np.float64)
        assert [int(raw[0]), int(raw[1])] == p.expectation(d)
        out[f"exp_map_{i}"] = d
        out[f"exp_raw_{i}"] = raw
        out[f"exp_int_{i}"] = np.array(p.expectation(d), dtype=np.int64)
    # one training step of the reference as written (train.py:18-26,35): model.forward(img).double() -> nn.BCELoss -> backward
    model = reference_loader.build_reference_model(0, K, 64, 96)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 64, 96, generator=g)
    uv = np.array([[[10, 20], [30, 40], [50, 10], [90, 60]], [[5, 5], [60, 30], [20, 50], [80, 8]]], dtype=np.float64)
    gt = torch.stack([ds.gauss_2d_batch(96, 64, 8, torch.tensor(uv[b, :, 0]), torch.tensor(uv[b, :, 1])) for b in range(2)])
    model.train()
    loss = torch.nn.BCELoss()(model.forward(x).double(), gt)
    loss.backward()
    out["step_uv"] = uv
    out["step_loss"] = np.array(loss.item())
    named = dict(model.named_parameters())
    pre = "resnet.resnet34_8s."
    for short in ("conv1.weight", "bn1.weight", "layer1.0.conv1.weight", "layer2.0.downsample.0.weight", "layer3.0.bn1.bias",
                  "layer4.2.bn2.weight", "fc.bias"):
        out["step_grad_" + short] = named[pre + short].grad.numpy().copy()
    out["step_grad_fc.weight_live"] = named[pre + "fc.weight"].grad[:K].numpy().copy()
    out["step_grad_fc.weight_dead_abs_sum"] = np.array(named[pre + "fc.weight"].grad[K:].abs().sum().item())
    out["step_grad_l4_conv2_sub"] = named[pre + "layer4.2.conv2.weight"].grad[::16, ::16].numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "golden_v2.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
