"""Thin object wrapper over the C-ABI context (`wb_ctx`, include/watsor_b200.h).

Host code stays Python, as in the reference; every numeric step happens inside
libwatsor_b200.so.  numpy arrays are only argument carriers (frames in, rows out).
"""
import ctypes
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_uint32, c_void_p, cast

import numpy as np

from . import _lib
from ._lib import ClassFilter, check
from .stream.share import MAX_DETECTIONS, Detection

# 0: fp32 CUDA-core convs; 1: bf16 tcgen05; 2: fp32 storage, dense convs as 3xTF32 tcgen05 (fp32-faithful)
PRECISION_FP32, PRECISION_BF16_TC, PRECISION_TF32X3 = 0, 1, 2


def _ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def _addr(obj):
    """Address of a numpy array / ctypes object / raw integer (device pointer)."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        assert obj.flags['C_CONTIGUOUS'], 'arrays handed to libwatsor_b200 must be C-contiguous'
        return obj.ctypes.data
    return ctypes.addressof(obj)


class Engine:
    """One `wb_ctx`: a model resident on one B200 plus per-camera filter tables."""

    def __init__(self, model_blob, device=0, max_batch=8, precision=PRECISION_FP32):
        self.lib = _lib.load()
        self._blob = bytes(model_blob)       # keep alive during wb_create
        self._ctx = c_void_p()
        check(self.lib.wb_create(device, self._blob, len(self._blob), max_batch, precision,
                                 byref(self._ctx)))
        self.device = device
        self.max_batch = max_batch
        self.precision = precision
        ih, iw, nc, na, nl = (c_int32() for _ in range(5))
        check(self.lib.wb_model_info(self._ctx, byref(ih), byref(iw), byref(nc), byref(na), byref(nl)))
        self.input_h, self.input_w = ih.value, iw.value
        self.num_classes, self.num_anchors, self.num_layers = nc.value, na.value, nl.value
        self.cameras = {}

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if self._ctx:
            self.lib.wb_destroy(self._ctx)
            self._ctx = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_name(self):
        buf = ctypes.create_string_buffer(255)
        check(self.lib.wb_device_name(self._ctx, buf, 255))
        return buf.value.decode()

    def set_stream(self, cuda_stream):
        check(self.lib.wb_set_stream(self._ctx, int(cuda_stream)))

    # ------------------------------------------------------------------- cameras
    def set_camera(self, cam_id, width, height, zone_rasters=None, class_filters=(), flags=0):
        """class_filters: iterable of (label, confidence, area, zones or None).  label -1 = default."""
        arr = (ClassFilter * max(1, len(class_filters)))()
        for i, (label, conf, area, zones) in enumerate(class_filters):
            bits = 0
            for z in (zones or ()):
                bits |= 1 << (z - 1)
            arr[i] = ClassFilter(label, 1 if zones else 0, bits, 0, conf, area)
        n_zones = 0
        raster_ptr = None
        dummy = np.zeros(1, np.uint8)          # stays alive until wb_set_camera has returned
        if zone_rasters is not None:
            zone_rasters = np.ascontiguousarray(zone_rasters, dtype=np.uint8)
            assert zone_rasters.ndim == 3 and zone_rasters.shape[1:] == (height, width)
            n_zones = zone_rasters.shape[0]
            # a mask without any zone is still "has_mask" (MaskFilter would reject everything)
            raster_ptr = zone_rasters.ctypes.data if n_zones else dummy.ctypes.data
        check(self.lib.wb_set_camera(self._ctx, cam_id, width, height, n_zones, raster_ptr,
                                     len(class_filters), arr, flags))
        self.cameras[cam_id] = (width, height)

    def register_host(self, address, nbytes):
        check(self.lib.wb_register_host(self._ctx, address, nbytes))

    def unregister_host(self, address):
        check(self.lib.wb_unregister_host(self._ctx, address))

    # ------------------------------------------------------------------ hot path
    def _io(self, frames, cam_ids, out, verdicts):
        n = len(frames)
        assert n == len(cam_ids)
        fp = _ptr_array([_addr(f) for f in frames])
        cams = (c_int32 * n)(*cam_ids)
        op = _ptr_array([_addr(o) for o in out]) if out is not None else None
        vp = _ptr_array([_addr(v) for v in verdicts]) if verdicts is not None else None
        return n, fp, cams, op, vp

    def detect(self, frames, cam_ids, out, verdicts=None, flags=0):
        """frames: host uint8 arrays (or device pointers with WB_F_FRAMES_ON_DEVICE);
        out: per frame a `Detection*100` ctypes array / address.  Returns device ms."""
        n, fp, cams, op, vp = self._io(frames, cam_ids, out, verdicts)
        ms = c_float(0)
        check(self.lib.wb_detect(self._ctx, n, fp, cams, flags, op, vp, byref(ms)))
        return ms.value

    def submit(self, slot, frames, cam_ids, flags=0):
        n, fp, cams, _, _ = self._io(frames, cam_ids, None, None)
        check(self.lib.wb_submit(self._ctx, slot, n, fp, cams, flags))

    def collect(self, slot, out=None, verdicts=None):
        op = _ptr_array([_addr(o) for o in out]) if out is not None else None
        vp = _ptr_array([_addr(v) for v in verdicts]) if verdicts is not None else None
        ms = c_float(0)
        check(self.lib.wb_collect(self._ctx, slot, op, vp, byref(ms)))
        return ms.value

    def stream_fence(self, cuda_stream, direction):
        """direction 0: slot streams wait for `cuda_stream`; 1: `cuda_stream` waits for the slots."""
        check(self.lib.wb_stream_fence(self._ctx, int(cuda_stream), direction))

    # --------------------------------------------------------------- frame scatter (NCCL behind the C-ABI)
    def comm_unique_id(self):
        """128-byte rendezvous id; the root rank makes it and ships it to the others over any host channel."""
        buf = (ctypes.c_uint8 * 128)()
        check(self.lib.wb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        """Collective over the `world` engines (one per process / GPU)."""
        assert len(unique_id) == 128
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.wb_comm_init(self._ctx, rank, world, buf))

    def scatter_frames(self, root, send_ptrs, recv_ptr, nbytes, cuda_stream=0):
        """send_ptrs: on the root, one device pointer per rank (that rank's `nbytes` slab), else None; recv_ptr:
        this rank's device buffer.  cuda_stream 0: later submits on this engine are ordered after the scatter."""
        sp = _ptr_array(send_ptrs) if send_ptrs is not None else None
        check(self.lib.wb_scatter_frames(self._ctx, root, sp, c_void_p(int(recv_ptr)), nbytes, int(cuda_stream)))

    def comm_destroy(self):
        check(self.lib.wb_comm_destroy(self._ctx))

    # --------------------------------------------------------------- stage level
    def preprocess(self, frames):
        n = len(frames)
        fp = _ptr_array([_addr(np.ascontiguousarray(f)) for f in frames])
        w = (c_int32 * n)(*[f.shape[1] for f in frames])
        h = (c_int32 * n)(*[f.shape[0] for f in frames])
        out = np.empty((n, self.input_h, self.input_w, 3), np.float32)
        check(self.lib.wb_preprocess(self._ctx, n, fp, w, h, out.ctypes.data))
        return out

    def backbone(self, pre, stop_layer=-1, layer_shape=None):
        pre = np.ascontiguousarray(pre, dtype=np.float32)
        n = pre.shape[0]
        enc = np.empty((n, self.num_anchors, 4), np.float32)
        logits = np.empty((n, self.num_anchors, self.num_classes + 1), np.float32)
        layer_out = None
        if stop_layer >= 0 and layer_shape is not None:
            layer_out = np.empty((n,) + tuple(layer_shape), np.float32)
        check(self.lib.wb_backbone(self._ctx, n, pre.ctypes.data, enc.ctypes.data, logits.ctypes.data,
                                   stop_layer, layer_out.ctypes.data if layer_out is not None else None,
                                   layer_out.size if layer_out is not None else 0))
        return enc, logits, layer_out

    def postprocess(self, enc, logits, cam_ids, flags=0):
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        logits = np.ascontiguousarray(logits, dtype=np.float32)
        n = enc.shape[0]
        rows = [(Detection * MAX_DETECTIONS)() for _ in range(n)]
        verd = np.zeros((n, MAX_DETECTIONS), np.uint32)
        boxes = np.empty((n, MAX_DETECTIONS, 4), np.float32)
        scores = np.empty((n, MAX_DETECTIONS), np.float32)
        classes = np.empty((n, MAX_DETECTIONS), np.float32)
        num = np.empty(n, np.int32)
        check(self.lib.wb_postprocess(
            self._ctx, n, enc.ctypes.data, logits.ctypes.data, (c_int32 * n)(*cam_ids), flags,
            _ptr_array([ctypes.addressof(r) for r in rows]),
            _ptr_array([verd[i].ctypes.data for i in range(n)]), boxes.ctypes.data, scores.ctypes.data,
            classes.ctypes.data, num.ctypes.data))
        return rows, verd, boxes, scores, classes, num

    def filter_rows(self, cam_id, rows, n_rows=None):
        """rows: ctypes Detection array (updated in place).  Returns uint32 verdicts."""
        n = n_rows if n_rows is not None else len(rows)
        verd = np.zeros(n, np.uint32)
        check(self.lib.wb_filter_rows(self._ctx, cam_id, n, ctypes.addressof(rows), verd.ctypes.data))
        return verd

    def anchors(self):
        out = np.empty((self.num_anchors, 4), np.float32)
        check(self.lib.wb_anchors(self._ctx, out.ctypes.data))
        return out

    def last_launch_count(self):
        n = c_int(0)
        check(self.lib.wb_last_launch_count(self._ctx, byref(n)))
        return n.value

    def profile_layers(self, device_frames, cam_ids):
        n = len(device_frames)
        cap = self.num_layers + 8
        ms = (c_float * cap)()
        kinds = (c_int32 * cap)()
        cnt = c_int(0)
        check(self.lib.wb_profile_layers(self._ctx, n, _ptr_array(list(device_frames)),
                                         (c_int32 * n)(*cam_ids), ms, kinds, cap, byref(cnt)))
        return [(kinds[i], ms[i]) for i in range(cnt.value)]
