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

CocoClass = namedtuple('CocoClass', ['index', 'label'])


def get_coco_class(idx):
    """Label record for an index; unknown indices map to 'unlabeled' (coco.py:124-131).  The reference's record
    also carries drawing attributes (box / font colours) for its output stage, which is out of scope here."""
    if 0 <= idx < len(COCO_CLASSES):
        return CocoClass(idx, COCO_CLASSES[idx])
    return CocoClass(0, COCO_CLASSES[0])
