"""hulk_keypoints_b200: B200-native (sm_100a) KeypointsGauss heatmap inference / training hot path.

Public surface mirrors the reference (vainaviv/hulk-keypoints):
    KeypointsGauss, Prediction, KeypointsDataset, gauss_2d_batch, transform
All computation goes through libhulk_sm100.so (include/hulk_sm100.h); there is no CPU fallback.
"""
from .model import KeypointsGauss  # noqa: F401
from .prediction import Prediction  # noqa: F401
from .dataset import KeypointsDataset, gauss_2d_batch, gauss_2d_targets, transform  # noqa: F401

__all__ = ["KeypointsGauss", "Prediction", "KeypointsDataset", "gauss_2d_batch", "gauss_2d_targets", "transform"]
