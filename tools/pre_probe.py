#!/usr/bin/env python
"""Runs the stand-alone resize+normalise kernel (wb_preprocess) on batches of 640x480 and 1920x1080 frames, for
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_preprocess`."""
import sys

import numpy as np

sys.path.insert(0, '.')
from tests.workload import v2_coco_model  # noqa: E402
from watsor_b200.engine import Engine  # noqa: E402

eng = Engine(v2_coco_model().to_blob(), device=0, max_batch=8)
rng = np.random.default_rng(0)
for (w, h) in ((640, 480), (1920, 1080)):
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(8)]
    for _ in range(3):
        out = eng.preprocess(frames)
    print(w, h, float(out.mean()))
