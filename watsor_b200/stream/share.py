"""The shared-memory frame ABI of watsor/stream/share.py:11-73, byte for byte.

Only the *data layout* of the stream runtime is part of the detection hot path
(SURVEY.md section 8b): `BoundingBox` 16 B, `Detection` 72 B, `Header` 7224 B and the
RGB24 image array that live in `multiprocessing.sharedctypes` memory.  The worker
skeletons, latches and queues of watsor/stream stay the reference's own.
`libwatsor_b200.so` writes `Detection[100]` blocks in exactly this layout
(include/watsor_b200.h: wb_detection).
"""
from ctypes import Structure, addressof, c_double, c_int, memset, sizeof
from multiprocessing import RLock
from multiprocessing.sharedctypes import Array, Value

from numpy import frombuffer

MAX_DETECTIONS = 100
MAX_ZONES = 10


class BoundingBox(Structure):
    _fields_ = [('x_min', c_int), ('y_min', c_int), ('x_max', c_int), ('y_max', c_int)]


class Detection(Structure):
    _fields_ = [('label', c_int), ('zones', c_int * MAX_ZONES), ('confidence', c_double),
                ('bounding_box', BoundingBox)]


class Header(Structure):
    _fields_ = [('width', c_int), ('height', c_int), ('channels', c_int), ('epoch', c_double),
                ('detections', Detection * MAX_DETECTIONS)]


assert sizeof(BoundingBox) == 16 and sizeof(Detection) == 72 and sizeof(Header) == 7224
assert Detection.zones.offset == 4 and Detection.confidence.offset == 48
assert Detection.bounding_box.offset == 56 and Header.detections.offset == 24


class Frame(object):
    """One frame slot: header + image under one lock (share.py:35-73).  The state latch of
    the reference (sync.py:79-103) is owned by the reference's stream runtime; a `latch`
    attribute can be attached by the caller."""

    def __init__(self, width, height, channels=3, array_type_code='B'):
        self.__lock = RLock()
        self.__header = Value(Header, width, height, channels, 0, lock=self.__lock)
        self.__image = Array(array_type_code, width * height * channels, lock=self.__lock)
        self.latch = None

    def clear(self):
        self.__header.epoch = 0
        memset(addressof(self.__image.get_obj()), 0, sizeof(self.__image.get_obj()))
        memset(addressof(self.__header.detections), 0, sizeof(self.__header.detections))

    @property
    def lock(self):
        return self.__lock

    @property
    def header(self):
        return self.__header

    @property
    def image(self):
        return self.__image

    def get_numpy_image(self, dtype=None):
        shape = (self.header.height, self.header.width, self.header.channels)
        return shape, frombuffer(self.image.get_obj(), dtype).reshape(shape)

    @property
    def image_address(self):
        return addressof(self.__image.get_obj())

    @property
    def image_nbytes(self):
        return sizeof(self.__image.get_obj())

    @property
    def detections_address(self):
        return addressof(self.__header.detections)


class FrameBuffer(object):
    """Ring of frames of one camera (share.py:76-81)."""

    def __init__(self, maxsize, width, height, channels=3, array_type_code='B'):
        self.__frames = [Frame(width, height, channels, array_type_code) for _ in range(maxsize)]

    @property
    def frames(self):
        return self.__frames
