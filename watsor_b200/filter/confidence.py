"""ConfidenceFilter -- same constructor and call contract as watsor/filter/confidence.py:5-19
(`confidence is not None and detection.confidence >= confidence`, threshold per label =
percent/100); the comparison itself runs in CUDA (csrc/kernels_post.cu: apply_filters)."""
from .. import _lib
from ..config.coco import COCO_CLASSES
from ._gpu import GpuPredicate


class ConfidenceFilter(GpuPredicate):
    verdict_bit = _lib.WB_V_CONFIDENCE

    def __init__(self, camera_config):
        self.thresholds = {}
        for entry in camera_config['detect']:
            coco_class = next(iter(entry))
            self.thresholds[COCO_CLASSES.index(coco_class)] = entry[coco_class]['confidence'] / 100
        super().__init__(camera_config.get('width', 1), camera_config.get('height', 1),
                         [(idx, thr, 0.0, None) for idx, thr in self.thresholds.items()])
