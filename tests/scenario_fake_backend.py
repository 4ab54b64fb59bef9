#!/usr/bin/env python
"""CPU scenario for tests/test_reference_runtime_worker.py: the drop-in ObjectDetector on the reference's OWN
stream runtime (Spin / Work / StateLatch / FrameBuffer from `watsor.stream`, PYTHONPATH supplied by the test) with
a fake pipelined back-end (submit / collect).  Checks the base-class set-up (`Spin._run`: thread name), that every
payload's latch advances exactly once and only after its rows were written, and the drain on stop."""
import json
import os
import sys
import time
from queue import Queue
from threading import Event, Thread, current_thread

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import watsor.stream.work as ref_work                      # noqa: E402
from watsor.stream.share import FrameBuffer                # noqa: E402
from watsor.stream.sync import State                       # noqa: E402

from watsor_b200.detection import detector as det_mod      # noqa: E402

assert det_mod.Work is ref_work.Work
events = []


class FakePipelined:
    max_batch = 3
    device_name = 'FAKE-PIPE:0'

    def __init__(self, model_path, device):
        self.slots = {}
        self.thread_name = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def configure_camera(self, cam, w, h, cfg):
        events.append(('configure', cam, w, h))

    def register_frame_buffer(self, fb):
        events.append(('register', len(fb.frames)))

    def submit(self, slot, images, cams, fuse_filters=False):
        assert slot not in self.slots, 'slot re-used before collect'
        self.thread_name = current_thread().name
        self.slots[slot] = [int(img[0, 0, 0]) for img in images]
        events.append(('submit', slot, len(images)))

    def collect(self, slot, rows):
        vals = self.slots.pop(slot)
        if 250 in vals:
            raise RuntimeError('boom')
        for v, r in zip(vals, rows):
            r[0].label = v + 1
        events.append(('collect', slot, len(vals)))
        return 0.75


def main():
    stop, q = Event(), Queue()
    fb = FrameBuffer(8, 16, 8)
    for i, f in enumerate(fb.frames):
        f.get_numpy_image(np.uint8)[1][0, 0, 0] = i
        f.latch.next()                                  # READY -> DETECT (1 party): as ReadDetectPublish would
    fb.frames[6].get_numpy_image(np.uint8)[1][0, 0, 0] = 250
    for i in range(8):
        q.put(ref_work.Payload('cam', i))
    w = det_mod.ObjectDetector(Thread, 'detector1', stop, Queue(), q, {'cam': fb},
                               kwargs={'detector_class': FakePipelined, 'detector_args': ('/m', 0)})
    w.start()
    t0 = time.time()
    while time.time() - t0 < 5 and not q.empty():
        time.sleep(0.01)
    time.sleep(0.3)
    stop.set()
    w.join(5)
    states = [int(f.latch.state) for f in fb.frames]
    labels = [f.header.detections[0].label for f in fb.frames]
    print(json.dumps({'states': states, 'labels': labels, 'events': events, 'publish': int(State.PUBLISH),
                      'fps': w.fps(), 'inference_time': w.inference_time(),
                      'device_name': w.device_name.decode(), 'alive': w.is_alive()}))


if __name__ == '__main__':
    main()
