"""Shared inputs of the visual-effects tests: random frames and Detection rows that exercise every placement of the
label (above / below / inside the box), labels cut by the right border, overlapping boxes, unknown label indices,
padding rows (label 0) and zones."""
import numpy as np

from watsor_b200.stream.share import MAX_DETECTIONS, Detection


def random_rows(rng, width, height, n_drawn, n_zones=0):
    rows = (Detection * MAX_DETECTIONS)()
    slots = sorted(rng.choice(MAX_DETECTIONS, size=n_drawn, replace=False).tolist())
    for k, i in enumerate(slots):
        d = rows[i]
        kind = k % 6
        x0 = int(rng.integers(0, width - 1))
        y0 = int(rng.integers(0, height - 1))
        if kind == 0:      # near the top: the label goes below or inside
            y0 = int(rng.integers(0, min(22, height - 1)))
        elif kind == 1:    # tall box from the top to the bottom: label inside
            y0 = int(rng.integers(0, 10))
        elif kind == 2:    # at the right border: the label is cut
            x0 = int(rng.integers(max(0, width - 60), width - 1))
        x1 = int(rng.integers(x0, width))
        y1 = int(rng.integers(y0, height))
        if kind == 1:
            y1 = height - 1 - int(rng.integers(0, 10))
        d.bounding_box.x_min, d.bounding_box.y_min, d.bounding_box.x_max, d.bounding_box.y_max = x0, y0, x1, y1
        d.label = int(rng.integers(1, 91)) if k % 11 else int(rng.integers(91, 200))
        d.confidence = float(np.float32(rng.random())) if k % 5 else [0.125, 0.005, 1.0, 0.995, 0.0][(k // 5) % 5]
        if n_zones and k % 2 == 0:
            zs = rng.choice(np.arange(1, n_zones + 1), size=int(rng.integers(1, min(n_zones, 3) + 1)), replace=False)
            for j, z in enumerate(zs):
                d.zones[j] = int(z)
    return rows


def random_alpha(rng, width, height, n_zones):
    """An alpha channel like a watsor mask: `n_zones` fully opaque blobs (the zones: filter/mask.py:84-88 finds the
    outlines of the alpha == 255 regions) on a translucent background of assorted alpha levels."""
    import cv2
    alpha = rng.integers(0, 200, (height // 8 + 1, width // 8 + 1)).astype(np.uint8)
    alpha = np.kron(alpha, np.ones((8, 8), np.uint8))[:height, :width].copy()
    for z in range(n_zones):
        cx = int((z + 0.5) * width / n_zones)
        cy = int(rng.integers(height // 4, 3 * height // 4))
        ax = int(rng.integers(6, max(7, width // (2 * n_zones) - 4)))
        ay = int(rng.integers(8, max(9, height // 5)))
        if z % 2:
            cv2.ellipse(alpha, (cx, cy), (ax, ay), 0.0, 0, 360, 255, -1)
        else:
            pts = np.array([[cx - ax, cy - ay], [cx + ax, cy - ay // 2], [cx + ax // 2, cy + ay], [cx - ax, cy + ay // 2]])
            cv2.fillPoly(alpha, [pts.astype(np.int32)], 255)
    return alpha
