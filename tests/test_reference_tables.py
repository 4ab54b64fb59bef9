"""Small tables and containers the filters and the frame ABI depend on, against the reference modules imported
from the read-only tree: the COCO label table (watsor/config/coco.py:14-131) and the Frame / FrameBuffer views
(watsor/stream/share.py:37-113).  CPU only; skipped where /root/reference is absent."""
import ctypes
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref_modules():
    sys.path.insert(0, REF)
    try:
        import watsor.config.coco as coco
        import watsor.stream.share as share
        yield coco, share
    finally:
        sys.path.remove(REF)


def test_coco_table_equals_reference(ref_modules):
    ref_coco, _ = ref_modules
    from watsor_b200.config import coco
    assert coco.COCO_CLASSES == list(ref_coco.COCO_CLASSES)
    for idx in list(range(-3, 95)) + [1000]:
        try:
            theirs = ref_coco.get_coco_class(idx)
        except Exception as e:                                  # whatever the reference does out of range ...
            with pytest.raises(type(e)):                         # ... we do the same
                coco.get_coco_class(idx)
            continue
        # the reference record also carries drawing attributes (colours, font) of the out-of-scope output stage
        assert coco.get_coco_class(idx).label == theirs.label, idx


def test_frame_views_equal_reference(ref_modules):
    _, ref_share = ref_modules
    from watsor_b200.stream import share
    for w, h in ((64, 48), (1, 1), (320, 240)):
        ours, theirs = share.Frame(w, h), ref_share.Frame(w, h, 3, 'B')
        so, io = ours.get_numpy_image(np.uint8)
        st, it = theirs.get_numpy_image(np.uint8)
        assert so == st == (h, w, 3) and io.shape == it.shape and io.dtype == it.dtype
        assert ctypes.sizeof(ours.header.get_obj()) == ctypes.sizeof(theirs.header.get_obj()) == 7224
        assert (ours.header.width, ours.header.height, ours.header.channels) == \
               (theirs.header.width, theirs.header.height, theirs.header.channels)
        io[...] = 7
        ours.header.detections[99].label = 5
        ours.clear()
        assert not io.any() and ours.header.detections[99].label == 0
    fb_o, fb_t = share.FrameBuffer(3, 32, 16), ref_share.FrameBuffer(3, 32, 16)
    assert len(fb_o.frames) == len(fb_t.frames) == 3
