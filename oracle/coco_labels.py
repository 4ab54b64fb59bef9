"""COCO-2014 "paper" label map (91 ids, 0 = unlabeled), the table the reference
indexes with `COCO_CLASSES.index(label)` (watsor/config/coco.py:14-105, used by
watsor/filter/confidence.py:14, area.py:14, mask.py:32).  TEST INFRASTRUCTURE.
Kept separate from the product's table on purpose: tests compare the two.
"""
COCO_CLASSES = (
    "unlabeled|person|bicycle|car|motorcycle|airplane|bus|train|truck|boat|traffic light|"
    "fire hydrant|street sign|stop sign|parking meter|bench|bird|cat|dog|horse|sheep|cow|"
    "elephant|bear|zebra|giraffe|hat|backpack|umbrella|shoe|eye glasses|handbag|tie|suitcase|"
    "frisbee|skis|snowboard|sports ball|kite|baseball bat|baseball glove|skateboard|surfboard|"
    "tennis racket|bottle|plate|wine glass|cup|fork|knife|spoon|bowl|banana|apple|sandwich|"
    "orange|broccoli|carrot|hot dog|pizza|donut|cake|chair|couch|potted plant|bed|mirror|"
    "dining table|window|desk|toilet|door|tv|laptop|mouse|remote|keyboard|cell phone|"
    "microwave|oven|toaster|sink|refrigerator|blender|book|clock|vase|scissors|teddy bear|"
    "hair drier|toothbrush"
).split("|")
assert len(COCO_CLASSES) == 91
