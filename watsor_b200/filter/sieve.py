"""The per-frame body of `DetectionSieve._incoming_frame` (watsor/filter/sieve.py:21-52): clone the
100 detection rows, run the filter chain, write the survivors back and zero-fill the rest.  The
worker class itself is a `WorkPassthroughPublish` of the reference's stream runtime and is only
defined when that runtime is importable (i.e. when this package is dropped into watsor)."""
from ctypes import addressof, memmove, memset, sizeof

from ..stream.share import Detection


def sieve_frame(detections, filters):
    """detections: the frame's `Detection * 100` array (modified in place).  Returns the OR of the
    filters' suspicious-activity flags.  The standard chain `[TrackFilter([...])]` (main.py:293-299) of this
    package's filters takes the native route: one CUDA call for the predicates, one C++ call for tracker +
    write-back (`TrackFilter.sieve`)."""
    if len(filters) == 1 and getattr(filters[0], 'can_sieve', False):
        return filters[0].sieve(detections)
    cloned = []
    for d in detections:
        c = Detection()
        memmove(addressof(c), addressof(d), sizeof(d))
        cloned.append(c)
    suspicious_activity = False
    current = cloned
    for flt in filters:
        current, sa = flt(current)
        suspicious_activity |= sa
    it = iter(current)
    for dst in detections:
        src = next(it, None)
        if src is not None:
            memmove(addressof(dst), addressof(src), sizeof(src))
        else:
            memset(addressof(dst), 0, sizeof(dst))
    return suspicious_activity


try:
    from watsor.stream.share import FramesPerSecond
    from watsor.stream.work import WorkPassthroughPublish

    class DetectionSieve(WorkPassthroughPublish):
        def __init__(self, name, stop_event, log_queue, frame_queue, frame_buffer, filters, decoder_rate_limiter,
                     kwargs=None):
            self.__fps = FramesPerSecond()
            super().__init__(name, stop_event, log_queue, frame_queue, frame_buffer,
                             args=(filters, decoder_rate_limiter, self.__fps), kwargs={} if kwargs is None else kwargs)

        @property
        def fps(self):
            return self.__fps

        def _incoming_frame(self, frame, stop_event, filters, decoder_rate_limiter, fps, *args, **kwargs):
            if sieve_frame(frame.header.detections, filters):
                if decoder_rate_limiter.unlimited():
                    self._logger.debug("FPS is unlimited due to an object detected")
            fps(value=True)
except ImportError:
    pass
