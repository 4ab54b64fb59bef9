"""AreaFilter -- contract of watsor/filter/area.py:5-26: threshold per label =
percent/100 * area(BoundingBox(0, 0, W-1, H-1)); passes when
abs((x_max-x_min+1)*(y_max-y_min+1)) >= threshold.  Evaluated in CUDA."""
from .. import _lib
from ..config.coco import COCO_CLASSES
from ._gpu import NEG_INF, GpuPredicate


class AreaFilter(GpuPredicate):
    verdict_bit = _lib.WB_V_AREA

    def __init__(self, camera_config):
        self.thresholds = {}
        width, height = camera_config['width'], camera_config['height']
        max_area = abs(((width - 1) - 0 + 1) * ((height - 1) - 0 + 1))
        for entry in camera_config['detect']:
            coco_class = next(iter(entry))
            self.thresholds[COCO_CLASSES.index(coco_class)] = entry[coco_class]['area'] / 100 * max_area
        super().__init__(width, height, [(idx, NEG_INF, thr, None) for idx, thr in self.thresholds.items()])
