"""keypoints_b200 — B200-native unsupervised-keypoint training path (drop-in for the hot path of
DuaneNielsen/keypoints: keypoints.models.{knn,vgg,keynet,transporter,functional}, tps, data_augments)."""
from . import config as _config
from .config import get_precision, set_precision, precision  # noqa: F401

__all__ = ['get_precision', 'set_precision', 'precision']
