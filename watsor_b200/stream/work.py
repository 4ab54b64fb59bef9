"""Minimal stand-ins for the pieces of the reference's stream runtime the detector worker is built
on: `Payload` (watsor/stream/work.py:42), a `Work`-shaped base class (spin.py:8-91 + work.py:9-33:
thread-or-process delegate, `_run` -> `_spin(_process)` loop, queue get with a 1 s timeout) and the
shared fps / inference-time counters read by `/metrics` (share.py:164-238).

When this package is dropped into a watsor checkout the reference's own classes are used instead
(`watsor_b200.detection.detector` imports `watsor.stream.*` first); these exist so that the batched
worker can be exercised -- and unit-tested on CPU -- without the rest of watsor.
"""
import logging
from collections import namedtuple
from multiprocessing import Value
from queue import Empty
from time import time

Payload = namedtuple('Payload', ['sender', 'frame_index'])


class RateCounter(object):
    """events per second over a sliding window (share.py:164-222 semantics, process-shared scalar)."""

    def __init__(self, timeframe=10.0):
        self._count = Value('i', 0)
        self._start = Value('d', 0.0)
        self._rate = Value('d', 0.0)
        self._timeframe = timeframe

    def __call__(self, value=None):
        now = time()
        with self._count.get_lock():
            if value is not None:
                if self._start.value == 0.0:
                    self._start.value = now
                self._count.value += 1
            elapsed = now - self._start.value if self._start.value else 0.0
            if elapsed >= self._timeframe:
                self._rate.value = self._count.value / elapsed
                self._count.value = 0
                self._start.value = now
            elif elapsed > 0 and self._rate.value == 0.0:
                return self._count.value / elapsed
            return self._rate.value


class MeanCounter(object):
    """running mean of the reported values (InferenceTime, share.py:225-238)."""

    def __init__(self):
        self._sum = Value('d', 0.0)
        self._n = Value('i', 0)

    def __call__(self, value=None):
        with self._sum.get_lock():
            if value is not None:
                self._sum.value += value
                self._n.value += 1
                if self._n.value > 100:
                    self._sum.value *= 0.5
                    self._n.value //= 2
            return self._sum.value / self._n.value if self._n.value else 0.0


class Work(object):
    """Same calling conventions as the reference: the delegate runs
    `_run(stop_event, log_queue, frame_queue, *args)`, `_spin(action, stop_event, *a)` calls
    `action(*a)`, `_process(frame_queue, *args)` hands a payload to `_next_frame(payload, *args)`."""

    def __init__(self, delegate_class, name, stop_event, log_queue, frame_queue, args=(), kwargs=None):
        self._logger = None
        self._delegate_class = delegate_class
        self._name = name
        self._stop_event = stop_event
        self._args = (stop_event, log_queue, frame_queue) + tuple(args)
        self._kwargs = dict(kwargs or {})
        self._delegate = None
        self.initialize()

    def initialize(self):
        self._delegate = self._delegate_class(name=self._name, target=self._run, args=self._args,
                                              kwargs=self._kwargs)

    @property
    def name(self):
        return self._name

    def start(self):
        self._delegate.start()

    def terminate(self):
        self._stop_event.set()

    def join(self, timeout=None):
        self._delegate.join(timeout)

    def is_alive(self):
        return self._delegate.is_alive()

    @staticmethod
    def _spin(action, stop_event, *args, **kwargs):
        while not stop_event.is_set():
            action(*args, **kwargs)

    def _config_logger(self, log_queue, *args, **kwargs):
        if self._logger is None:
            self._logger = logging.getLogger(self.__class__.__name__)

    def _run(self, stop_event, log_queue, *args, **kwargs):
        self._config_logger(log_queue, *args, **kwargs)
        self._spin(self._process, stop_event, *args, **kwargs)

    def _process(self, frame_queue, *args, **kwargs):
        try:
            payload = frame_queue.get(timeout=1)
        except Empty:
            return
        if payload is not None:
            self._next_frame(payload, *args, **kwargs)

    def _next_frame(self, *args, **kwargs):
        pass
