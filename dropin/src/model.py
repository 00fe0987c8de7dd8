"""Drop-in for reference src/model.py: re-export the B200-native KeypointsGauss."""
from hulk_keypoints_b200 import KeypointsGauss  # noqa: F401
