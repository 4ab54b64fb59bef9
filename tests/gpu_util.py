"""Helpers shared by the `-m gpu` parity tests."""
import ctypes

import numpy as np

from watsor_b200.stream.share import MAX_DETECTIONS, Detection


def rows_to_tuples(rows, n=MAX_DETECTIONS):
    return [(rows[r].label, rows[r].confidence, rows[r].bounding_box.x_min, rows[r].bounding_box.y_min,
             rows[r].bounding_box.x_max, rows[r].bounding_box.y_max) for r in range(n)]


def zones_of(rows, n=MAX_DETECTIONS):
    return [list(rows[r].zones) for r in range(n)]


def new_rows(count=1):
    return [(Detection * MAX_DETECTIONS)() for _ in range(count)]


def rows_bytes(rows):
    return bytes(ctypes.string_at(ctypes.addressof(rows), ctypes.sizeof(rows)))


def compare_rows(got, want, boxes64=None, image_shape=None, conf_tol=1e-3, margin_px=2e-3):
    """The parity bar of BASELINE.json's north star: class ids exact, confidences within 1e-3,
    integer box coordinates exact -- except where the float64 evaluation of the same graph puts
    the coordinate within `margin_px` of an integer boundary (a genuine rounding tie: fp32
    summation order decides it).  Returns the number of such tolerated +-1 flips."""
    flips = 0
    assert len(got) == len(want)
    for r, (g, w) in enumerate(zip(got, want)):
        assert g[0] == w[0], ('label', r, g, w)
        assert abs(g[1] - w[1]) <= conf_tol, ('confidence', r, g, w)
        for k in range(4):
            if g[2 + k] == w[2 + k]:
                continue
            assert boxes64 is not None and r < len(boxes64), ('box', r, g, w)
            # rows: x_min,y_min,x_max,y_max  <- boxes: ymin,xmin,ymax,xmax
            coord = boxes64[r][[1, 0, 3, 2][k]]
            scale = (image_shape[1] - 1) if k in (0, 2) else (image_shape[0] - 1)
            v = min(max(coord, 0.0), 1.0) * scale
            assert abs(g[2 + k] - w[2 + k]) == 1 and abs(v - round(v)) <= margin_px, ('box', r, k, g, w, v)
            flips += 1
    return flips
