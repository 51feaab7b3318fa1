"""Rendering parity (keras_inference.py:137-174) without a GPU: integer box geometry, line thickness, class colours (the
full 80-entry colormap of tools/utils.py:89-105) and label placement against the reference's own rendered result."""
import json
import os

import numpy as np

from conftest import GOLDEN
from k210_yolo_framework_b200.draw import box_geometry, draw_detections
from k210_yolo_framework_b200.helper import Helper, _COLORMAP


def test_colormap_is_the_references_80_entries():
    assert len(_COLORMAP) == 80
    assert _COLORMAP[0] == (255, 82, 0) and _COLORMAP[19] == (255, 0, 245)          # the 20 VOC colours
    assert _COLORMAP[20] == (128, 0, 0) and _COLORMAP[40] == (61, 230, 250) and _COLORMAP[79] == (11, 200, 200)
    h = Helper(None, 80, np.zeros((2, 3, 2)), np.array([[224, 320]] * 2), np.array([[7, 10], [14, 20]]))
    assert h.colormap is _COLORMAP


def test_box_geometry_rounding_and_clipping():
    # floor(v + 0.5) rounding, clipped to the image; the label origin is (left, top + 1) for every box inside the image
    assert box_geometry((224, 320), (25.9, 188.63, 74.67, 309.19), (28, 7)) == (26, 189, 75, 309, (189, 27))
    assert box_geometry((224, 320), (-3.2, -1.0, 230.4, 400.0), (28, 7)) == (0, 0, 224, 320, (0, 1))
    assert box_geometry((224, 320), (224.0, 10.0, 300.0, 20.0), (28, 7))[4] == (10, 224 - 7)   # the reference's (never-met) branch


def test_render_matches_reference_result_image(dog_u8, dog_golden):
    found = [tuple(d) for d in dog_golden["keras"]["detections"]]
    img = np.array(draw_detections(dog_u8, found, _COLORMAP, font_path="does-not-exist.otf"))
    with open(os.path.join(GOLDEN, "dog_res_labels.json")) as fh:
        ref = json.load(fh)
    assert (224 + 320) // 300 == 1                                                  # one-pixel outlines on this image
    for c, _idx, _score, top, left, bottom, right in found:
        colour = np.array(_COLORMAP[int(c)])
        t, l, b, r, origin = box_geometry((224, 320), (top, left, bottom, right), (0, 0))
        hit = (img == colour).all(-1)
        assert hit[t, l:r + 1].all() and hit[b, l:r + 1].all() and hit[t:b + 1, l].all() and hit[t:b + 1, r].all()   # the rectangle
        # the filled label box starts where the reference's does (its right / bottom edge depends on the font file)
        g = ref[str(int(c))]
        ys, xs = np.nonzero(hit[origin[1]:origin[1] + 6, origin[0]:origin[0] + 12])
        assert len(ys) > 20                                                          # filled, not just an outline
        assert abs(origin[1] - g["ymin"]) <= 1 and abs(origin[0] - g["xmin"]) <= 2   # JPEG ringing allows a pixel or two
