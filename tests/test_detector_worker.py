"""Host logic of the batched detector worker (watsor_b200/detection/detector.py) on CPU, with a fake
back-end: drains up to max_batch payloads per tick, one latch count-down per payload (also when the
back-end raises), metrics updated, reference-protocol back-ends still served one by one."""
import ctypes
import time
from queue import Queue
from threading import Event, Thread

import numpy as np

from watsor_b200.detection import detector as det_mod
from watsor_b200.detection.detector import ObjectDetector
from watsor_b200.stream.share import FrameBuffer
from watsor_b200.stream.work import Payload


class Latch:
    def __init__(self):
        self.count = 0

    def next(self):
        self.count += 1


class FakeBatched:
    max_batch = 4
    batches = []
    configured = []

    def __init__(self, model_path, device):
        self.device = device

    device_name = 'FAKE:0'

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def configure_camera(self, cam, w, h, cfg):
        FakeBatched.configured.append((cam, w, h, cfg))

    def detect_batch(self, images, cams, rows, fuse_filters=False):
        FakeBatched.batches.append(len(images))
        for img, r in zip(images, rows):
            if img[0, 0, 0] == 255:
                raise RuntimeError('boom')
            r[0].label = int(img[0, 0, 0]) + 1
        return 1.5


class FakeSingle(FakeBatched):
    detect_batch = None

    def detect(self, shape, image, rows):
        rows[0].label = 7
        return 2.0


def run_worker(detector_class, frames_to_send, fb, extra=None):
    stop, q = Event(), Queue()
    for i in frames_to_send:
        q.put(Payload('cam', i))
    kwargs = {'detector_class': detector_class, 'detector_args': ('/model', 0)}
    kwargs.update(extra or {})
    w = ObjectDetector(Thread, 'detector1', stop, Queue(), q, {'cam': fb}, kwargs=kwargs)
    w.start()
    t0 = time.time()
    while time.time() - t0 < 5 and sum(f.latch.count for f in fb.frames) < len(frames_to_send):
        time.sleep(0.01)
    stop.set()
    w.join(3)
    return w


def make_fb(n=6):
    fb = FrameBuffer(n, 16, 8)
    for i, f in enumerate(fb.frames):
        f.latch = Latch()
        f.get_numpy_image(np.uint8)[1][0, 0, 0] = i
    return fb


def test_batched_drain_and_latch():
    FakeBatched.batches, FakeBatched.configured = [], []
    fb = make_fb()
    w = run_worker(FakeBatched, range(6), fb, {'camera_configs': {'cam': {'detect': []}}})
    assert [f.latch.count for f in fb.frames] == [1] * 6
    assert [f.header.detections[0].label for f in fb.frames] == [1, 2, 3, 4, 5, 6]
    assert sum(FakeBatched.batches) == 6 and max(FakeBatched.batches) <= 4 and len(FakeBatched.batches) <= 3
    assert FakeBatched.configured == [(0, 16, 8, {'detect': []})]          # once per camera
    assert w.device_name == b'FAKE:0' and w.inference_time() == 1.5 and w.fps() >= 0.0


def test_latch_advances_when_backend_raises():
    FakeBatched.batches = []
    fb = make_fb(2)
    fb.frames[1].get_numpy_image(np.uint8)[1][0, 0, 0] = 255
    run_worker(FakeBatched, [1], fb)
    assert fb.frames[1].latch.count == 1                                   # detector.py:111-112 `finally`


def test_reference_protocol_backend_is_served_per_frame():
    fb = make_fb(3)
    run_worker(FakeSingle, range(3), fb)
    assert [f.header.detections[0].label for f in fb.frames] == [7, 7, 7]
    assert [f.latch.count for f in fb.frames] == [1, 1, 1]


def test_factory_needs_model_and_gpu(tmp_path):
    import pytest
    with pytest.raises(AssertionError, match='Failed to create an object detector'):
        det_mod.create_object_detectors(Thread, Event(), Queue(), Queue(), {}, str(tmp_path))
