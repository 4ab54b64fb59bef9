"""`create_object_detectors` / `ObjectDetector` -- the factory and worker the application calls
(watsor/main.py:414-418), re-implemented for a batched accelerator.  Same signature, same
`device_name` / `fps` / `inference_time` attributes (read by `/metrics`, main.py:242-251), same
latch protocol (exactly one `frame.latch.next()` per drained payload, also on failure --
detector.py:111-112), back-end constructed inside the child so that CUDA state is created after
`spawn` (detector.py:84-96).

Difference from watsor/detection/detector.py:102-112: the worker drains every payload that is
already waiting (at most one per camera: `BalancedQueue` holds a 1-slot semaphore per camera,
sync.py:156-166) and hands them to the B200 as ONE batch; the confidence / area / mask predicates
of the camera run fused behind the NMS when camera configs are supplied via
`kwargs['camera_configs']`.
"""
from multiprocessing.sharedctypes import Array
from os import path
from queue import Empty

from numpy import uint8

try:  # dropped into a watsor checkout: build on the reference's own runtime
    from watsor.stream.share import FramesPerSecond, InferenceTime
    from watsor.stream.work import Payload, Work
except ImportError:  # stand-alone
    from ..stream.work import MeanCounter as InferenceTime
    from ..stream.work import Payload, Work
    from ..stream.work import RateCounter as FramesPerSecond

from .b200 import COMPILED_MODEL, MODEL_FILES
from .devices import b200_gpus


def has_model(model_path):
    return any(path.isfile(path.join(model_path, f)) for f in (COMPILED_MODEL,) + MODEL_FILES)


def create_object_detectors(delegate_class, stop_event, log_queue, frame_queue, frame_buffers, model_path, kwargs=None):
    """One batched detector worker per visible B200 (detector.py:12-55).  There is deliberately no
    CPU fallback here: without a GPU or a model the assertion below fires, as in the reference."""
    detectors = []
    kwargs = {} if kwargs is None else kwargs
    if has_model(model_path):
        for device, clazz in b200_gpus():
            name = 'detector{}'.format(len(detectors) + 1)
            detectors.append(ObjectDetector(delegate_class, name, stop_event, log_queue, frame_queue, frame_buffers,
                                            kwargs={**kwargs, 'detector_class': clazz,
                                                    'detector_args': (model_path, device)}))
    assert len(detectors) > 0, "Failed to create an object detector. " \
                               "Make sure a B200 is visible and model files are provided."
    return detectors


class ObjectDetector(Work):
    def __init__(self, delegate_class, name, stop_event, log_queue, frame_queue_in, frame_buffers, kwargs=None):
        self.__fps = FramesPerSecond()
        self.__inference_time = InferenceTime()
        self.__device_name = Array('c', 255)
        super().__init__(delegate_class, name, stop_event, log_queue, frame_queue_in,
                         args=(stop_event, frame_buffers, self.__fps, self.__inference_time),
                         kwargs={} if kwargs is None else kwargs)

    @property
    def device_name(self):
        return self.__device_name.value

    @property
    def fps(self):
        return self.__fps

    @property
    def inference_time(self):
        return self.__inference_time

    def _run(self, stop_event, log_queue, *args, **kwargs):
        # logging set-up of the base runtime (reference: Spin._run / _config_logger, spin.py:51-73)
        self._config_logger(log_queue, *args, **kwargs)
        try:
            detector_class = kwargs.get('detector_class')
            detector_args = kwargs.get('detector_args')
            with detector_class(*detector_args) as object_detector:
                self.__device_name.value = str.encode(object_detector.device_name)[:len(self.__device_name) - 1]
                self._logger.debug("{}{} initialized".format(object_detector.__class__.__name__, detector_args))
                self._cameras = {}
                self._spin(self._process, stop_event, *args, object_detector, **kwargs)
        except FileNotFoundError as e:
            self._logger.error(e)
        except Exception:
            self._logger.exception('Detection failure')

    # -- one tick: block for the first payload, then take whatever else is already queued
    def _process(self, frame_queue, *args, **kwargs):
        try:
            first = frame_queue.get(timeout=1)
        except Empty:
            return
        if first is None:
            return
        payloads = [first]
        object_detector = args[-1]
        limit = getattr(object_detector, 'max_batch', 1)
        while len(payloads) < limit:
            try:
                nxt = frame_queue.get_nowait()
            except Empty:
                break
            if nxt is not None:
                payloads.append(nxt)
        self._next_frames(payloads, *args, **kwargs)

    def _next_frame(self, payload, *args, **kwargs):         # single-payload entry of the base class
        self._next_frames([payload], *args, **kwargs)

    def _next_frames(self, payloads, stop_event, frame_buffers, fps, inference_time, object_detector, *args, **kwargs):
        frames = [frame_buffers[p.sender].frames[p.frame_index] for p in payloads]
        try:
            if callable(getattr(object_detector, 'detect_batch', None)):
                images, cams, rows = [], [], []
                for p, frame in zip(payloads, frames):
                    shape, image_np = frame.get_numpy_image(uint8)
                    images.append(image_np)
                    cams.append(self._camera_id(object_detector, p.sender, shape, kwargs.get('camera_configs')))
                    rows.append(frame.header.detections)
                # fuse_filters stays off: the sieve thread applies (and zone-marks) the predicates itself
                ms = object_detector.detect_batch(images, cams, rows, fuse_filters=False)
                for _ in payloads:
                    inference_time(value=ms)
                    fps(value=True)
            else:                                            # any reference-protocol back-end
                for frame in frames:
                    shape, image_np = frame.get_numpy_image(uint8)
                    inference_time(value=object_detector.detect(shape, image_np, frame.header.detections))
                    fps(value=True)
        finally:
            for frame in frames:
                if getattr(frame, 'latch', None) is not None:
                    frame.latch.next()

    def _camera_id(self, object_detector, sender, shape, camera_configs):
        cam = self._cameras.get(sender)
        if cam is None:
            cam = len(self._cameras)
            cfg = (camera_configs or {}).get(sender)
            object_detector.configure_camera(cam, int(shape[1]), int(shape[0]), cfg)
            self._cameras[sender] = cam
        return cam
