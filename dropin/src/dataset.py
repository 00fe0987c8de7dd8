"""Drop-in for reference src/dataset.py."""
from hulk_keypoints_b200 import KeypointsDataset, gauss_2d_batch, transform  # noqa: F401
