"""Thread safety of the drop-in filter path: the reference runs one DetectionSieve thread per camera in the main
process (ref: watsor/main.py:378-384) and ctypes releases the GIL, so `wb_filter_rows` / `wb_set_camera` of the
process-wide filter engine are called concurrently."""
import threading

import numpy as np
import pytest

from tests import workload
from tests.gpu_util import rows_to_tuples, zones_of
from watsor_b200.stream.share import Detection

pytestmark = pytest.mark.gpu


def random_rows(rng, n=100):
    rows = (Detection * n)()
    for r in range(n):
        rows[r].label = int(rng.integers(0, 91))
        rows[r].confidence = float(rng.random())
        x0, x1 = sorted(int(v) for v in rng.integers(0, 640, 2))
        y0, y1 = sorted(int(v) for v in rng.integers(0, 480, 2))
        rows[r].bounding_box.x_min, rows[r].bounding_box.x_max = x0, x1
        rows[r].bounding_box.y_min, rows[r].bounding_box.y_max = y0, y1
    return rows


def test_four_camera_threads_hammer_the_filter_engine():
    """4 cameras with different masks, each judged by its own thread 200 times; every call must return exactly
    what the same call returns single-threaded (rows of one camera must never be judged against another camera's
    table or overwritten by another thread's staging copy)."""
    from watsor_b200.filter.area import AreaFilter
    from watsor_b200.filter.confidence import ConfidenceFilter
    from watsor_b200.filter.mask import MaskFilter
    from watsor_b200.filter.track import TrackFilter
    import ctypes
    cams = [0, 1, 2, 3]
    filters, inputs, expected = {}, {}, {}
    for c in cams:
        cfg = workload.camera_config(c)
        filters[c] = TrackFilter([ConfidenceFilter(cfg), AreaFilter(cfg), MaskFilter(cfg)])
        rng = np.random.default_rng(100 + c)
        inputs[c] = [random_rows(rng) for _ in range(5)]
    from watsor_b200.filter._gpu import filter_engine

    def judge(c, rows):
        filters[c]._ensure_slot()
        work = (Detection * 100)()
        ctypes.memmove(ctypes.addressof(work), ctypes.addressof(rows), ctypes.sizeof(rows))
        v = filter_engine().filter_rows(filters[c]._slot, work, 100)
        return [int(x) for x in v], zones_of(work), rows_to_tuples(work)

    for c in cams:
        expected[c] = [judge(c, rows) for rows in inputs[c]]
    assert len({str(expected[c][0][0]) for c in cams}) > 1          # the cameras really judge differently
    errors = []

    def worker(c):
        try:
            for it in range(200):
                k = it % 5
                if judge(c, inputs[c][k]) != expected[c][k]:
                    errors.append((c, it))
                    return
        except Exception as e:              # noqa
            errors.append((c, repr(e)))

    threads = [threading.Thread(target=worker, args=(c,)) for c in cams]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def test_set_camera_while_other_threads_filter():
    """Lazy binding of a new camera's table (wb_set_camera: device sync + table upload) while other cameras'
    threads keep filtering."""
    from watsor_b200.filter.confidence import ConfidenceFilter
    from watsor_b200.filter.track import TrackFilter
    cfg = workload.camera_config(1, mask=False)
    base = TrackFilter([ConfidenceFilter(cfg)])
    rng = np.random.default_rng(7)
    rows = random_rows(rng)
    want = [d.label > 0 and d.confidence >= 0.5 for d in rows]
    stop = threading.Event()
    bad = []

    def spin():
        while not stop.is_set():
            got = base._passing(list(rows))
            if len(got) != sum(want):
                bad.append(len(got))
                return

    t = threading.Thread(target=spin)
    t.start()
    for i in range(20):
        f = TrackFilter([ConfidenceFilter(workload.camera_config(i % 8, mask=False))])
        assert len(f._passing(list(rows))) == sum(want)
    stop.set()
    t.join()
    assert not bad
