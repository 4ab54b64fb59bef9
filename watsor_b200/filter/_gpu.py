"""Process-wide filter engine: a model-less `wb_ctx` whose camera slots hold the tables of
stand-alone filter objects, so that `ConfidenceFilter(cfg)(detection)` & co. run the same
CUDA predicate code (`apply_filters`, csrc/kernels_post.cu) as the fused detector path."""
import ctypes
import threading

import numpy as np

from .. import _lib
from ..engine import Engine
from ..model import Model
from ..stream.share import Detection

_lock = threading.Lock()
_engine = None
_free_slots = None
NEG_INF = float('-inf')


def _null_model_blob():
    m = Model(name='filters-only', input_h=16, input_w=16, num_classes=1, num_anchors=1)
    m.anchors_tensor = m.add_tensor(np.zeros((1, 4), np.float32))
    return m.to_blob()


def filter_engine(device=0):
    global _engine, _free_slots
    with _lock:
        if _engine is None:
            _engine = Engine(_null_model_blob(), device=device, max_batch=1)
            _free_slots = list(range(255, -1, -1))
        return _engine


def alloc_slot():
    filter_engine()
    with _lock:
        if not _free_slots:
            raise _lib.WatsorB200Error('all 256 stand-alone filter slots are in use')
        return _free_slots.pop()


def free_slot(slot):
    with _lock:
        if _free_slots is not None and slot is not None:
            _free_slots.append(slot)


class GpuPredicate(object):
    """Base of the filter objects: a table of (label, confidence, area, zones) rows, optional
    zone rasters, and the verdict bit that decides `__call__`."""

    verdict_bit = _lib.WB_V_PASS

    def __init__(self, width, height, class_filters, zone_rasters=None):
        self.width, self.height = int(width), int(height)
        self.class_filters = list(class_filters)
        self.zone_rasters = zone_rasters
        self._slot = None

    def _bind(self):
        if self._slot is None:
            self._slot = alloc_slot()
            filter_engine().set_camera(self._slot, self.width, self.height, self.zone_rasters,
                                       self.class_filters, flags=_lib.WB_CAM_NO_LABEL_CHECK)
        return self._slot

    def __call__(self, detection):
        slot = self._bind()
        row = (Detection * 1)()
        ctypes.memmove(ctypes.addressof(row), ctypes.addressof(detection), ctypes.sizeof(Detection))
        verdict = filter_engine().filter_rows(slot, row, 1)[0]
        for z in range(len(detection.zones)):
            detection.zones[z] = row[0].zones[z]
        return bool(verdict & self.verdict_bit)

    def __del__(self):
        try:
            free_slot(self._slot)
        except Exception:
            pass
