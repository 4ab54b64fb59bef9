"""COCO label table: index <-> label, as used by the filter configuration
(`COCO_CLASSES.index(label)` in watsor/filter/confidence.py:14, area.py:14, mask.py:32;
table in watsor/config/coco.py:14-105).  Ids follow the COCO-2014 "paper" numbering that
TensorFlow Object-Detection models emit (1 = person ... 90 = toothbrush; 0 = unlabeled)."""
from collections import namedtuple

_THINGS = """person bicycle car motorcycle airplane bus train truck boat traffic_light fire_hydrant
street_sign stop_sign parking_meter bench bird cat dog horse sheep cow elephant bear zebra giraffe hat
backpack umbrella shoe eye_glasses handbag tie suitcase frisbee skis snowboard sports_ball kite
baseball_bat baseball_glove skateboard surfboard tennis_racket bottle plate wine_glass cup fork knife
spoon bowl banana apple sandwich orange broccoli carrot hot_dog pizza donut cake chair couch
potted_plant bed mirror dining_table window desk toilet door tv laptop mouse remote keyboard cell_phone
microwave oven toaster sink refrigerator blender book clock vase scissors teddy_bear hair_drier
toothbrush"""

COCO_CLASSES = ['unlabeled'] + [w.replace('_', ' ') for w in _THINGS.split()]
assert len(COCO_CLASSES) == 91

CocoClass = namedtuple('CocoClass', ['index', 'label', 'box_color', 'font_color', 'box_thickness', 'font_thickness',
                                     'font_scale', 'alpha'])

# Drawing attributes of the output stage (watsor/config/coco.py:107-121): one random RGB colour per label from a fixed
# seed, white text, 1-pixel box and strokes, font scale 0.5, label box opacity 0.55.
_BOX_COLORS = None


def _box_colors():
    global _BOX_COLORS
    if _BOX_COLORS is None:
        import numpy as np
        _BOX_COLORS = np.random.RandomState(255).uniform(0, 255, size=(len(COCO_CLASSES), 3)).astype(np.uint8)
    return _BOX_COLORS


def get_coco_class(idx):
    """Record for a label index; unknown indices map to 'unlabeled' (coco.py:124-131)."""
    if not 0 <= idx < len(COCO_CLASSES):
        idx = 0
    return CocoClass(idx, COCO_CLASSES[idx], tuple(int(v) for v in _box_colors()[idx]), (255, 255, 255), 1, 1, 0.5, 0.55)
