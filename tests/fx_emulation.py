"""numpy emulation of k_fx_prepare / k_fx_render (watsor_b200/csrc/kernels_fx.cu): the same geometry, tables and
order of operations, applied per detection instead of per pixel.  Lets the CPU suite check the ALGORITHM against the
oracle (OpenCV) without a GPU; the GPU tests then only have to show that the kernel equals it."""
import numpy as np

from watsor_b200.config.coco import COCO_CLASSES, get_coco_class


def add_weighted_lut(color, alpha):
    """cv2.addWeighted(u8, alpha, solid colour, 1 - alpha, 0): rint(fma(a, alpha, b * beta)) in float32."""
    fa, fb = np.float32(alpha), np.float32(1.0 - alpha)
    lut = np.empty((3, 256), np.uint8)
    a = np.arange(256, dtype=np.float64)
    for c in range(3):
        t = np.float32(np.float32(color[c]) * fb)                       # one rounding
        r = (a * np.float64(fa) + np.float64(t)).astype(np.float32)    # exact product + sum, one rounding: the FMA
        lut[c] = np.clip(np.rint(r), 0, 255).astype(np.uint8)
    return lut


def percent_digits(confidence):
    pct = np.rint(float(confidence) * 100.0)
    n = int(pct) if 0.0 <= pct < 1.0e6 else 0
    return str(n)


def render(atlas, image_in, detections, alpha_channel=None, contours=None, blend=True, draw=True, outlines=True):
    h, w = image_in.shape[:2]
    out = image_in.copy()
    if blend and alpha_channel is not None:
        af = (alpha_channel.astype(np.float32) / np.float32(255)).astype(np.float32)
        wi = (np.float32(255) * (np.float32(1) - af)).astype(np.float32)
        acc = (image_in.astype(np.float32) * af[:, :, None]).astype(np.float32) + wi[:, :, None]
        out = acc.astype(np.float32).astype(np.int32).astype(np.uint8)
    zone_sel = 0
    margin = int(round(np.ceil(0.1 * atlas.text_height)))
    th, bl = atlas.text_height, atlas.baseline
    if draw:
        for d in detections:
            if not d.label > 0:
                continue
            bb = d.bounding_box
            left, top, right, bottom = bb.x_min, bb.y_min, bb.x_max, bb.y_max
            style = d.label if d.label < len(COCO_CLASSES) else 0
            cls = get_coco_class(style)
            x0, x1, y0, y1 = min(left, right), max(left, right), min(top, bottom), max(top, bottom)
            ys, xs = np.mgrid[0:h, 0:w]
            border = (((ys == y0) | (ys == y1)) & (xs >= x0) & (xs <= x1)) | \
                     (((xs == x0) | (xs == x1)) & (ys >= y0) & (ys <= y1))
            out[border] = cls.box_color
            text = cls.label + ': ' + percent_digits(d.confidence) + '%'
            pen = sum(atlas.advance[c] for c in text)
            text_w = int(np.rint(pen * 0.5 + 1.0))
            total = th + 2 * margin
            if top - bl > total:
                text_bottom = top
            elif bottom + total + bl < h:
                text_bottom = bottom + total + bl
            else:
                text_bottom = top + total + bl
            p1x, p1y = left, text_bottom - bl - th - 2 * margin
            p2x, p2y = left + text_w + 2 * margin, text_bottom
            if p1x >= 0 and p1y >= 0 and min(p1y, h) < min(p2y, h):
                lut = add_weighted_lut(cls.box_color, cls.alpha)
                patch = out[min(p1y, h):min(p2y, h), min(p1x, w):min(p2x, w)]
                for c in range(3):
                    patch[:, :, c] = lut[c][patch[:, :, c]]
                atlas.draw(out, text, (left + margin, text_bottom - bl - margin))
            for z in d.zones:
                if 0 < z <= 32:
                    zone_sel |= 1 << (z - 1)
    if draw and outlines and contours is not None and zone_sel:
        out[(contours & np.uint32(zone_sel)) != 0] = (255, 255, 0)
    return out
