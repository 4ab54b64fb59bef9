"""Synthetic frames with known content: the reference's `Artist` recipe
(watsor/test/detect_stream.py:42-70) -- four quadrants, in each a random ellipse / rectangle
/ right triangle filled (200,200,200) on black.

Restated rather than imported: the original hands float bounds to `random.randrange`
(`left + width` with width = image.width / 2), which Python >= 3.12 rejects; with the
integral bounds used here the random stream and the drawing calls are the same as the
original's under the Python it was written for, for even frame sizes.
"""
import math
import random

import numpy as np
from PIL import Image, ImageDraw

SHAPE_TRIANGLE, SHAPE_ELLIPSE, SHAPE_RECTANGLE = 1, 2, 3


def _draw_random_shape(draw, left, top, width, height):
    fill = (200, 200, 200)
    shape = random.choice([SHAPE_ELLIPSE, SHAPE_RECTANGLE, SHAPE_TRIANGLE])
    p1_x = random.randrange(left, left + math.floor(width / 2))
    p1_y = random.randrange(top, top + math.floor(height / 2))
    p2_x = random.randrange(left + math.ceil(width / 2), left + width)
    p2_y = random.randrange(top + math.ceil(height / 2), top + height)
    if shape == SHAPE_ELLIPSE:
        draw.ellipse([(p1_x, p1_y), (p2_x, p2_y)], fill)
    elif shape == SHAPE_RECTANGLE:
        draw.rectangle([(p1_x, p1_y), (p2_x, p2_y)], fill)
    else:
        draw.polygon([(p1_x, p1_y), (p1_x, p2_y), (p2_x, p2_y)], fill)
    return shape, (p1_x, p1_y, p2_x, p2_y)


def draw_random_shapes(image, draw):
    width, height = image.width // 2, image.height // 2
    cx, cy = image.width // 2, image.height // 2
    return [_draw_random_shape(draw, 0, 0, width, height),
            _draw_random_shape(draw, cx, 0, width, height),
            _draw_random_shape(draw, cx, cy, width, height),
            _draw_random_shape(draw, 0, cy, width, height)]


def artist_frame(width, height, cam=0, frame=0, with_truth=False):
    random.seed(1000 * cam + frame)
    with Image.new('RGB', (width, height)) as image:
        truth = draw_random_shapes(image, ImageDraw.Draw(image))
        arr = np.array(image)
    return (arr, truth) if with_truth else arr
