"""TrackFilter -- contract of watsor/filter/track.py:8-149.

Stage 1 (track.py:26): keep detections with `label > 0` that pass every predicate.  When the
predicates are this package's Confidence/Area/Mask filters they are merged into ONE per-camera
table and all 100 rows are judged by a single CUDA call (`wb_filter_rows`), instead of up to 300
Python calls; zones are written only on rows whose earlier predicates passed, exactly like the
reference's lazy `all(...)`.  Foreign callables are still honoured (applied on the host, in order,
after the fused ones).

Stage 2 (track.py:29-149): the centroid tracker -- per label, match each known object (newest
first by nearest centroid), append to its history (deque, `history` long), drop unmatched objects,
start new ones, and report objects seen at least `sensitivity` times as the envelope of their
history.  Sequential, tiny and stateful: it stays on the host.
"""
import ctypes
from collections import OrderedDict, deque

import numpy as np

from .. import _lib
from ..stream.share import MAX_DETECTIONS, Detection
from ._gpu import NEG_INF, alloc_slot, filter_engine, free_slot
from .area import AreaFilter
from .confidence import ConfidenceFilter
from .mask import MaskFilter


def _clone(detection):
    c = Detection()
    ctypes.memmove(ctypes.addressof(c), ctypes.addressof(detection), ctypes.sizeof(Detection))
    return c


def _centroid(bb):
    return int((bb.x_min + bb.x_max) / 2.0), int((bb.y_min + bb.y_max) / 2.0)


class TrackFilter(object):
    def __init__(self, filters=None, sensitivity=5, history=10):
        self.sensitivity = sensitivity
        self.history = history
        self.filters = [] if filters is None else list(filters)
        self.objects = OrderedDict()                    # label -> list of deque(history)
        self._slot = None
        self._fused, self._foreign = self._merge_tables(self.filters)

    # ------------------------------------------------------------------ stage 1
    @staticmethod
    def _merge_tables(filters):
        """[ConfidenceFilter][AreaFilter][MaskFilter] (any subset, in that order) -> one table."""
        conf = next((f for f in filters if isinstance(f, ConfidenceFilter)), None)
        area = next((f for f in filters if isinstance(f, AreaFilter)), None)
        mask = next((f for f in filters if isinstance(f, MaskFilter)), None)
        ours = [f for f in (conf, area, mask) if f is not None]
        order_ok = [f for f in filters if f in ours] == ours
        foreign = [f for f in filters if f not in ours]
        if not ours or not order_ok:
            return None, list(filters)
        labels = None
        for f in (conf, area):
            if f is not None:
                keys = set(f.thresholds)
                labels = keys if labels is None else labels & keys
        zones_by_label = mask.zones_by_label if mask is not None else {}
        rows = []
        if labels is None:                              # mask only: every label passes the first two
            rows.append((-1, NEG_INF, 0.0, None))
            labels = set(zones_by_label)
        for lab in sorted(labels):
            rows.append((lab, conf.thresholds[lab] if conf else NEG_INF, area.thresholds[lab] if area else 0.0,
                         zones_by_label.get(lab)))
        geom = mask or area or conf
        return {'rows': rows, 'rasters': mask.zone_rasters if mask is not None else None,
                'width': geom.width, 'height': geom.height}, foreign

    def _passing(self, detections):
        if self._fused is None:
            return [d for d in detections if d.label > 0 and all(f(d) for f in self.filters)]
        if self._slot is None:
            self._slot = alloc_slot()
            filter_engine().set_camera(self._slot, self._fused['width'], self._fused['height'],
                                       self._fused['rasters'], self._fused['rows'])
        dets = list(detections)
        kept = []
        for base in range(0, len(dets), MAX_DETECTIONS):
            chunk = dets[base:base + MAX_DETECTIONS]
            rows = (Detection * len(chunk))()
            for i, d in enumerate(chunk):
                ctypes.memmove(ctypes.addressof(rows[i]), ctypes.addressof(d), ctypes.sizeof(Detection))
            verdicts = filter_engine().filter_rows(self._slot, rows, len(chunk))
            for i, d in enumerate(chunk):
                for z in range(len(d.zones)):
                    d.zones[z] = rows[i].zones[z]
                if verdicts[i] & _lib.WB_V_PASS and all(f(d) for f in self._foreign):
                    kept.append(d)
        return kept

    def __call__(self, detections):
        return self._group_and_update(self._passing(detections))

    def __del__(self):
        try:
            free_slot(self._slot)
        except Exception:
            pass

    # ------------------------------------------------------------------ stage 2
    def _group_and_update(self, detections):
        groups = OrderedDict()
        for d in detections:
            groups.setdefault(d.label, []).append(d)
        suspicious_activity = len(groups) > 0
        for label in [l for l in self.objects if l not in groups]:
            del self.objects[label]
        for label, dets in groups.items():
            known = self.objects.setdefault(label, [])
            new_c = np.array([_centroid(d.bounding_box) for d in dets], dtype=np.int64).reshape(-1, 2)
            old_c = np.array([_centroid(h[0].bounding_box) for h in known], dtype=np.int64).reshape(-1, 2)
            used_rows, used_cols = set(), set()
            if len(known) and len(dets):
                diff = old_c[:, None, :].astype(np.float64) - new_c[None, :, :].astype(np.float64)
                dist = np.sqrt((diff ** 2).sum(-1))
                rows = np.argsort(dist.min(axis=1))
                cols = dist.argmin(axis=1)[rows]
                for r, c in zip(rows, cols):
                    if r in used_rows or c in used_cols:
                        continue
                    known[r].append(dets[c])
                    used_rows.add(int(r))
                    used_cols.add(int(c))
            for r in sorted(set(range(len(old_c))) - used_rows, reverse=True):
                del known[r]
            for c in sorted(set(range(len(dets))) - used_cols):
                known.append(deque([dets[c]], maxlen=self.history))
        result = []
        for label, known in self.objects.items():
            for h in known:
                if len(h) >= self.sensitivity:
                    result.append(self._combine(h))
        return result, suspicious_activity

    @staticmethod
    def _combine(h):
        out = _clone(h[0])
        for d in list(h)[1:]:
            out.confidence = max(out.confidence, d.confidence)
            out.bounding_box.x_min = min(out.bounding_box.x_min, d.bounding_box.x_min)
            out.bounding_box.y_min = min(out.bounding_box.y_min, d.bounding_box.y_min)
            out.bounding_box.x_max = max(out.bounding_box.x_max, d.bounding_box.x_max)
            out.bounding_box.y_max = max(out.bounding_box.y_max, d.bounding_box.y_max)
        zones = set()
        for d in h:
            zones.update(z for z in d.zones if z > 0)
        it = iter(zones)
        for i in range(len(out.zones)):
            out.zones[i] = next(it, 0)
        return out
