# Same module-level names as the reference's config.py (consumed with `from config import *`).
NUM_KEYPOINTS = 4
IMG_HEIGHT = 480
IMG_WIDTH = 640
GAUSS_SIGMA = 8
epochs = 25
batch_size = 4
