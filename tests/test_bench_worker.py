"""bench.py's e2e_worker harness (watsor_b200/bench_worker.py) on CPU with a fake back-end: the worker really runs
in a spawned process, frames are in shared memory, payloads cross a multiprocessing.Queue."""
import argparse

import numpy as np

from tests.fake_backend import FakeB200
from watsor_b200.bench_worker import run_worker_bench
from watsor_b200.model import Model


def test_worker_harness_counts_every_frame_once():
    args = argparse.Namespace(cameras=4, model='fake', frames='random', precision='tf32x3', steps=50, min_seconds=0.0)

    def load_model(kind):
        m = Model(name='x', input_h=16, input_w=16, num_classes=1, num_anchors=1)
        m.anchors_tensor = m.add_tensor(np.zeros((1, 4), np.float32))
        return m, 'fake'

    def make_frames(kind, cam, count):
        return [np.full((24, 32, 3), 10 * cam + i, np.uint8) for i in range(count)]

    rec = run_worker_bench(args, 0, lambda cam, kind: None, load_model, make_frames, width=32, height=24,
                           detector_class=FakeB200)
    assert rec['value'] > 0 and rec['ticks'] >= 200 and rec['rows_written']
    assert rec['device_name'] == 'FAKE-B200:0' and abs(rec['worker_inference_ms'] - 0.2) < 0.02
    print(rec)
