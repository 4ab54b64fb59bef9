"""The synthetic workload of BASELINE.json configs[2] / configs[3] (SURVEY.md 8(d) recipes), shared by
bench.py and the parity tests so that what is measured is what is parity-tested:

  * 8 cameras of 640x480 RGB per GPU,
  * SSD-MobileNet-v2 300x300, 90 COCO classes, score threshold 1e-8 (the model-zoo export value),
    seeded synthetic weights (no v2 weights exist offline),
  * a mask per camera: camera 0 = the reference's config/porch.png (tests/golden/porch.png, 2 zones);
    cameras 1.. = synthetic RGBA masks (background alpha 216, 1 + cam % 4 filled rectangles / ellipses of
    alpha 255 drawn from default_rng(cam)),
  * per-class thresholds = the schema defaults confidence 50 / area 10 / zones []
    (ref: watsor/config/schema.py:87-105) for every COCO label.
"""
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORCH = os.path.join(ROOT, 'tests', 'golden', 'porch.png')
W, H = 640, 480
_mask_dir = None


def synthetic_mask_rgba(cam, width=W, height=H):
    """RGBA uint8 [H,W,4]; zone pixels have alpha 255 (ref: watsor/filter/mask.py:78-88 reads alpha == 255)."""
    import cv2
    rng = np.random.default_rng(cam)
    img = np.zeros((height, width, 4), np.uint8)
    img[..., :3] = 40
    img[..., 3] = 216
    k = 1 + cam % 4
    for i in range(k):
        # one shape per vertical band so that zones never merge into one contour
        x_lo, x_hi = width * i // k, width * (i + 1) // k
        cx = int(rng.integers(x_lo + (x_hi - x_lo) // 3, x_hi - (x_hi - x_lo) // 3))
        cy = int(rng.integers(height // 4, 3 * height // 4))
        rx = int(rng.integers(max(8, (x_hi - x_lo) // 6), max(9, (x_hi - x_lo) // 3)))
        ry = int(rng.integers(height // 10, height // 4))
        alpha = np.zeros((height, width), np.uint8)
        if rng.integers(0, 2):
            cv2.rectangle(alpha, (max(x_lo + 2, cx - rx), cy - ry), (min(x_hi - 3, cx + rx), cy + ry), 255, -1)
        else:
            cv2.ellipse(alpha, (cx, cy), (min(rx, cx - x_lo - 2, x_hi - 3 - cx), ry), 0, 0, 360, 255, -1)
        img[alpha == 255, 3] = 255
    return img


def mask_png(cam, width=W, height=H):
    """Path of the camera's mask file (written once per process to a temp directory; camera 0 of the 640x480
    workload is the reference's porch.png)."""
    global _mask_dir
    if cam % 8 == 0 and (width, height) == (W, H) and os.path.isfile(PORCH):
        return PORCH
    import cv2
    if _mask_dir is None:
        _mask_dir = tempfile.mkdtemp(prefix='wb200_masks_')
    p = os.path.join(_mask_dir, 'mask_cam%d_%dx%d.png' % (cam, width, height))
    if not os.path.isfile(p):
        rgba = synthetic_mask_rgba(cam, width, height)
        assert cv2.imwrite(p, rgba[..., [2, 1, 0, 3]])
    return p


def coco_detect_defaults(labels=None):
    """`detect:` list with the schema defaults for every label (schema.py:87-105)."""
    if labels is None:
        from watsor_b200.config.coco import COCO_CLASSES
        labels = COCO_CLASSES[1:]
    return [{label: {'confidence': 50, 'area': 10, 'zones': []}} for label in labels]


def camera_config(cam, width=W, height=H, labels=None, mask=True):
    cfg = {'width': width, 'height': height, 'detect': coco_detect_defaults(labels)}
    if mask:
        cfg['mask'] = mask_png(cam, width, height)
    return cfg


def v2_coco_model():
    """SSD-MobileNet-v2, 90 classes, threshold 1e-8, seeded synthetic weights (configs[1..3])."""
    from watsor_b200.model import synthetic_ssd_mobilenet_v2
    return synthetic_ssd_mobilenet_v2(num_classes=90, seed=0, score_thr=1e-8)
