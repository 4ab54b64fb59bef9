"""CPU restatement of the reference's visual effects -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows watsor/output/copy.py:13-18 (CopyImageEffect), watsor/output/blend.py:8-32 (BlendEffect) and
watsor/output/draw.py:9-103 (DrawEffect, DrawEffectWithContours), with numpy and the same OpenCV calls the reference
makes (cv2.rectangle, cv2.getTextSize, cv2.addWeighted, cv2.putText, cv2.drawContours): OpenCV is the reference's own
arithmetic for this path, so calling it here is not a restatement risk.  Pinned against the reference's own classes,
imported from the read-only tree and run on the same inputs, in tests/test_oracle_effects.py (CPU).

A detection is anything with `.label`, `.confidence`, `.zones` and `.bounding_box.{x_min,y_min,x_max,y_max}` (the
ctypes rows of watsor_b200/stream/share.py qualify).
"""
import cv2
import numpy as np

from watsor_b200.config.coco import get_coco_class
from watsor_b200.filter.mask import find_contours

ZONE_OUTLINE = (255, 255, 0)           # draw.py:103


def blend_tables(alpha_channel):
    """blend.py:15-22: float32 alpha factor per pixel and channel, and the white share 255 * (1 - factor)."""
    factor = alpha_channel[:, :, np.newaxis].astype(np.float32) / 255
    factor = np.repeat(factor, 3, axis=2)
    white = np.full(factor.shape, 255, np.float32)
    white *= (1 - factor)
    return factor, white


def blend(image_in, image_out, tables):
    """blend.py:27-32"""
    factor, white = tables
    acc = np.zeros(factor.shape, np.float32)
    np.copyto(acc, image_in, casting='safe')
    acc *= factor
    acc += white
    np.copyto(image_out, acc, casting='unsafe')


def label_text(detection):
    """draw.py:14-15"""
    cls = get_coco_class(detection.label)
    return '{}: {}'.format(cls.label, '{0:.0%}'.format(detection.confidence)), cls


def draw_one(image, height, box, text, cls):
    """draw.py:51-88 for one detection; box = (left, top, right, bottom)."""
    left, top, right, bottom = box
    cv2.rectangle(image, (left, top), (right, bottom), cls.box_color, cls.box_thickness)
    if not text:
        return
    face = cv2.FONT_HERSHEY_DUPLEX
    (text_w, text_h), baseline = cv2.getTextSize(text, face, cls.font_scale, cls.font_thickness)
    margin = int(round(np.ceil(0.1 * text_h)))
    band = text_h + 2 * margin
    if top - baseline > band:                      # above the box
        text_bottom = top
    elif bottom + band + baseline < height:        # below it
        text_bottom = bottom + band + baseline
    else:                                          # inside, at the top
        text_bottom = top + band + baseline
    y0, y1 = text_bottom - baseline - text_h - 2 * margin, text_bottom
    x0, x1 = left, left + text_w + 2 * margin
    patch = image[y0:y1, x0:x1]
    if len(patch) == 0:
        return
    solid = np.full(patch.shape, cls.box_color, dtype=np.uint8)
    if len(solid) == 0:
        return
    mixed = cv2.addWeighted(patch, cls.alpha, solid, 1 - cls.alpha, 0)
    if mixed is None:
        return
    image[y0:y1, x0:x1] = mixed
    cv2.putText(image, text, (left + margin, text_bottom - baseline - margin), face, cls.font_scale, cls.font_color,
                cls.font_thickness, cv2.LINE_AA)


def draw(image_out, shape, detections):
    """draw.py:11-23"""
    for d in detections:
        if d.label > 0:
            text, cls = label_text(d)
            bb = d.bounding_box
            draw_one(image_out, shape[0], (bb.x_min, bb.y_min, bb.x_max, bb.y_max), text, cls)


def draw_zone_outlines(image_out, contours, detections):
    """draw.py:100-103"""
    for d in detections:
        if d.label > 0:
            for z in d.zones:
                if z > 0:
                    cv2.drawContours(image_out, contours, z - 1, color=ZONE_OUTLINE, thickness=1)


def effect_chain(image_in, detections, alpha_channel=None, do_draw=True):
    """What VisualEffects produces in image_out for one frame with the chain of main.py:302-312 (header copy aside):
    with a mask BlendEffect + DrawEffectWithContours, without CopyImageEffect + DrawEffect."""
    out = np.empty_like(image_in)
    if alpha_channel is not None:
        blend(image_in, out, blend_tables(alpha_channel))
    else:
        np.copyto(out, image_in)
    if do_draw:
        draw(out, image_in.shape, detections)
        if alpha_channel is not None:
            draw_zone_outlines(out, find_contours(alpha_channel), detections)
    return out
