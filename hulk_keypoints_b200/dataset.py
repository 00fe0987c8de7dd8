"""Target generation and dataset -- host-side mirror of reference src/dataset.py.

`gauss_2d_batch(width, height, sigma, U, V, normalize_dist=False)` keeps the reference signature
(dataset.py:36-44) and result: a (K, H, W) float64 CUDA tensor whose values are the float32 Gaussian
`exp(-((x-u)^2+(y-v)^2)/(2*sigma^2))` widened to double.  The ~12 torch launches + 2 host->device grid
copies per sample of the reference become one launch of hk_gauss_targets.  `gauss_2d_targets` is the
batched form used by the training step.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops


def _to_tensor(img: np.ndarray) -> torch.Tensor:
    """HWC uint8 (as returned by cv2.imread, BGR) -> CHW float32 in [0,1]; what the reference's
    `transforms.Compose([transforms.ToTensor()])` does (dataset.py:16), without a torchvision import."""
    if img.ndim == 2:
        img = img[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
    return t.float().div(255) if t.dtype == torch.uint8 else t.float()


transform = _to_tensor


def gauss_2d_targets(uv: torch.Tensor, height: int, width: int, sigma: float,
                     dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """Batched targets: uv (B,K,2) = (x, y) on the GPU -> (B,K,H,W)."""
    return ops.gauss_targets(uv, height, width, sigma, dtype)


def gauss_2d_batch(width, height, sigma, U, V, normalize_dist=False):
    """Drop-in for reference dataset.py:36-44.  U, V: (K,) tensors (any dtype, any device)."""
    U = torch.as_tensor(U).reshape(-1)
    V = torch.as_tensor(V).reshape(-1)
    uv = torch.stack([U.float(), V.float()], dim=-1).unsqueeze(0).cuda()
    G32 = ops.gauss_targets(uv, int(height), int(width), float(sigma), torch.float32)[0]
    if normalize_dist:
        return ops.l1_normalize_dim1(G32)  # reference dataset.py:33-34,42-43: F.normalize(p=1) = L1 over dim 1 (H), then .double()
    return G32.double()


class KeypointsDataset(Dataset):
    """Drop-in for reference dataset.py:52-79: `%05d.jpg` images + `%05d.npy` (K,2)=(x,y) labels, clipped to
    the image; returns (img float32 (3,H,W) on the host, gaussians float64 (K,H,W) on the GPU)."""

    def __init__(self, img_folder, labels_folder, num_keypoints, img_height, img_width, transform, gauss_sigma=8):
        self.num_keypoints = num_keypoints
        self.img_height = img_height
        self.img_width = img_width
        self.gauss_sigma = gauss_sigma
        self.transform = transform
        self.imgs = []
        self.labels = []
        for i in range(len(os.listdir(labels_folder))):
            label = np.load(os.path.join(labels_folder, '%05d.npy' % i)).reshape(num_keypoints, 2)
            label[:, 0] = np.clip(label[:, 0], 0, self.img_width - 1)
            label[:, 1] = np.clip(label[:, 1], 0, self.img_height - 1)
            self.imgs.append(os.path.join(img_folder, '%05d.jpg' % i))
            self.labels.append(torch.from_numpy(label).cuda())

    def __getitem__(self, index):
        import cv2

        img = self.transform(cv2.imread(self.imgs[index]))
        labels = self.labels[index]
        gaussians = gauss_2d_batch(self.img_width, self.img_height, self.gauss_sigma, labels[:, 0], labels[:, 1])
        return img, gaussians

    def __len__(self):
        return len(self.labels)
