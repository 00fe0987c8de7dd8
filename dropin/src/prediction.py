"""Drop-in for reference src/prediction.py."""
from hulk_keypoints_b200 import Prediction  # noqa: F401
