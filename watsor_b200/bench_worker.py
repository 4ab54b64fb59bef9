"""End-to-end throughput through the drop-in detector WORKER (watsor_b200.detection.detector.ObjectDetector):
frames live in `multiprocessing` shared memory (`FrameBuffer`, the reference's frame ABI), payloads travel through
a `multiprocessing.Queue`, the worker runs in its own process under `spawn` exactly as watsor starts its detectors
(ref: watsor/main.py:414-418, watsor/detection/detector.py:12-55), Detection rows come back in the shared frame
headers and every frame's latch is advanced once.  bench.py reports the figure as `e2e_worker`.

The feeder stands in for the decoder threads: per tick it enqueues one payload per camera and keeps at most
`DEPTH` ticks outstanding.  Frame contents are static (decode is out of scope, as for `e2e`).
"""
import os
import tempfile
import time
import multiprocessing
from multiprocessing import get_context

import numpy as np

DEPTH = 6          # ticks in flight between feeder and worker (queue + the worker's pipeline slots)
RING = 8           # frames per camera (> DEPTH, so a frame is never re-used while it is being detected)


class CountingLatch(object):
    """Stand-in for the reference's StateLatch on the frames of this harness: `next()` is what the worker calls
    exactly once per payload; a process-shared counter tells the feeder how many frames are done."""

    def __init__(self, counter):
        self.counter = counter

    def next(self, *a):
        with self.counter.get_lock():
            self.counter.value += 1


def run_worker_bench(args, local_rank, camera_config, load_model, make_frames, width=640, height=480,
                     detector_class=None):
    from .detection.b200 import B200ObjectDetector
    from .detection.detector import ObjectDetector
    from .stream.share import FrameBuffer
    from .stream.work import Payload
    # the shared frames, counters and queues must come from the same start-method context as the worker process
    # (watsor calls set_start_method('spawn') in main.py:474 before it builds anything)
    prev_method = multiprocessing.get_start_method(allow_none=True)
    multiprocessing.set_start_method('spawn', force=True)
    try:
        return _run(args, local_rank, camera_config, load_model, make_frames, width, height, detector_class,
                    B200ObjectDetector, ObjectDetector, FrameBuffer, Payload)
    finally:
        multiprocessing.set_start_method(prev_method, force=True)


def _run(args, local_rank, camera_config, load_model, make_frames, width, height, detector_class, B200ObjectDetector,
         ObjectDetector, FrameBuffer, Payload):
    ctx = get_context('spawn')
    C = args.cameras
    model, _ = load_model(args.model)
    tmp = tempfile.mkdtemp(prefix='wb200_worker_model_')
    with open(os.path.join(tmp, 'b200.wb200'), 'wb') as f:
        f.write(model.to_blob())
    done = ctx.Value('q', 0)
    names = ['cam%d' % c for c in range(C)]
    buffers = {}
    for c, name in enumerate(names):
        fb = FrameBuffer(RING, width, height)
        imgs = make_frames(args.frames, c, RING)
        for frame, img in zip(fb.frames, imgs):
            np.copyto(frame.get_numpy_image(np.uint8)[1], img)
            frame.latch = CountingLatch(done)
        buffers[name] = fb
    stop, frame_queue, log_queue = ctx.Event(), ctx.Queue(), ctx.Queue()
    env_before = {k: os.environ.get(k) for k in ('WATSOR_B200_MAX_BATCH', 'WATSOR_B200_PRECISION')}
    os.environ['WATSOR_B200_MAX_BATCH'] = str(C)
    os.environ['WATSOR_B200_PRECISION'] = args.precision
    worker = ObjectDetector(ctx.Process, 'detector1', stop, log_queue, frame_queue, buffers,
                            kwargs={'detector_class': detector_class or B200ObjectDetector, 'detector_args': (tmp, local_rank),
                                    'camera_configs': {n: camera_config(c, args.model) for c, n in enumerate(names)}})
    worker.start()
    for k, v in env_before.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v

    sent = 0

    def feed(ticks, deadline):
        nonlocal sent
        for _ in range(ticks):
            while sent - done.value >= DEPTH * C:
                if time.perf_counter() > deadline:
                    raise TimeoutError('worker stalled: %d of %d frames done' % (done.value, sent))
                time.sleep(0)
            idx = (sent // C) % RING
            for name in names:
                frame_queue.put(Payload(name, idx))
            sent += C
        while done.value < sent:
            if time.perf_counter() > deadline:
                raise TimeoutError('worker stalled: %d of %d frames done' % (done.value, sent))
            time.sleep(0)

    try:
        feed(30, time.perf_counter() + 120.0)                    # start-up: CUDA context, model upload, graph capture
        ticks = int(min(20000, max(args.steps, 200)))
        t0 = time.perf_counter()
        feed(ticks, t0 + 120.0)
        dt = time.perf_counter() - t0
        # a second, longer pass when the first was short
        if dt < args.min_seconds:
            ticks = int(min(20000, ticks * args.min_seconds / max(dt, 1e-3) * 1.1))
            t0 = time.perf_counter()
            feed(ticks, t0 + 120.0)
            dt = time.perf_counter() - t0
        rec = {'value': C * ticks / dt, 'unit': 'frames/s', 'ticks': ticks, 'ms_per_tick': 1e3 * dt / ticks,
               'device_name': worker.device_name.decode(errors='replace'), 'worker_fps_metric': worker.fps(),
               'worker_inference_ms': worker.inference_time(),
               'api': 'watsor_b200.detection.detector.ObjectDetector in a spawned Process; frames in multiprocessing '
                      'shared memory (FrameBuffer, pinned by the worker), payloads through multiprocessing.Queue, '
                      '%d cameras, <= %d ticks outstanding, Detection rows written into the shared frame headers' % (C, DEPTH),
               'h2d_bytes_per_tick': C * width * height * 3, 'd2h_bytes_per_tick': C * 7200}
        labels = [buffers[names[0]].frames[i].header.detections[0].label for i in range(RING)]
        rec['rows_written'] = bool(all(l >= 1 for l in labels))
        return rec
    finally:
        stop.set()
        worker.join(20)
        if worker.is_alive():
            try:
                worker._delegate.terminate()
            except Exception:
                pass
