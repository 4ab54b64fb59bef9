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
history.  Sequential, tiny and stateful: it runs on the host, in C++ (`csrc/tracker.cpp`,
`wb_tracker_update`), which also pins down what the reference leaves to its
runtime (argsort ties, the iteration order of Python sets -- see the header of tracker.cpp).
`sieve()` is the whole DetectionSieve body for the standard chain `[TrackFilter([...])]`
(main.py:293-299) in two calls: `wb_filter_rows` (CUDA) + `wb_sieve_rows` (host).
"""
import ctypes
from ctypes import byref, c_int, c_uint32, c_void_p

from .. import _lib
from ..stream.share import MAX_DETECTIONS, Detection
from ._gpu import NEG_INF, alloc_slot, filter_engine, free_slot
from .area import AreaFilter
from .confidence import ConfidenceFilter
from .mask import MaskFilter


def _ok(rc, what):
    if rc != 0:
        raise _lib.WatsorB200Error('%s failed (status %d)' % (what, rc))


def _clone(detection):
    c = Detection()
    ctypes.memmove(ctypes.addressof(c), ctypes.addressof(detection), ctypes.sizeof(Detection))
    return c


class TrackFilter(object):
    def __init__(self, filters=None, sensitivity=5, history=10):
        self.sensitivity = sensitivity
        self.history = history
        self.filters = [] if filters is None else list(filters)
        self._slot = None
        self._tracker = c_void_p()
        _ok(_lib.load().wb_tracker_create(sensitivity, history, byref(self._tracker)), 'wb_tracker_create')
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

    def _ensure_slot(self):
        if self._slot is None:
            self._slot = alloc_slot()
            filter_engine().set_camera(self._slot, self._fused['width'], self._fused['height'],
                                       self._fused['rasters'], self._fused['rows'])

    def _passing(self, detections):
        if self._fused is None:
            return [d for d in detections if d.label > 0 and all(f(d) for f in self.filters)]
        self._ensure_slot()
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

    def sieve(self, rows):
        """DetectionSieve._incoming_frame for filters == [self] (sieve.py:21-52): `rows` (Detection * n,
        the frame's header rows) are judged, tracked and rewritten in place.  Only available when every
        predicate is one of this package's (no foreign callables)."""
        assert self.can_sieve
        n = len(rows)
        verdicts = None
        if self._fused is not None:
            self._ensure_slot()
            verdicts = (c_uint32 * n)()
            for base in range(0, n, MAX_DETECTIONS):
                m = min(MAX_DETECTIONS, n - base)
                chunk = (Detection * m).from_address(ctypes.addressof(rows) + base * ctypes.sizeof(Detection))
                v = filter_engine().filter_rows(self._slot, chunk, m)
                for i in range(m):
                    verdicts[base + i] = v[i]
        sa = c_int()
        _ok(_lib.load().wb_sieve_rows(self._tracker, rows, n, verdicts, byref(sa)), 'wb_sieve_rows')
        return bool(sa.value)

    @property
    def can_sieve(self):
        return not self._foreign

    def __del__(self):
        try:
            free_slot(self._slot)
        except Exception:
            pass
        try:
            if self._tracker:
                _lib.load().wb_tracker_destroy(self._tracker)
                self._tracker = c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ stage 2
    def _group_and_update(self, detections):
        dets = list(detections)
        n = len(dets)
        rows = (Detection * max(n, 1))()
        for i, d in enumerate(dets):
            ctypes.memmove(ctypes.addressof(rows[i]), ctypes.addressof(d), ctypes.sizeof(Detection))
        out = (Detection * max(n, 1))()
        n_out, sa = c_int(), c_int()
        _ok(_lib.load().wb_tracker_update(self._tracker, rows, n, None, out, n, byref(n_out), byref(sa)),
            'wb_tracker_update')
        return [_clone(out[i]) for i in range(n_out.value)], bool(sa.value)
