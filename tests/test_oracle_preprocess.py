"""Oracle restatement of the reference's letterbox pre-processing (tools/utils.py:357-406) — CPU only."""
import numpy as np

from oracle import preprocess_ref
from k210_yolo_framework_b200 import preprocess


def test_identity_when_image_has_network_size(dog_u8):
    """data/dog.jpg is 224x320: scale 1, translation 0 -> the warp is the identity (what the pinned known answers rely on)."""
    scale, tr, inv = preprocess_ref.letterbox_params(dog_u8.shape[:2], (224, 320))
    assert scale.tolist() == [1.0, 1.0] and tr.tolist() == [0, 0]
    np.testing.assert_array_equal(inv, np.eye(3))
    np.testing.assert_array_equal(preprocess_ref.letterbox(dog_u8, (224, 320)), dog_u8)


def test_params_match_reference_formula_on_people(people_u8):
    """374x499 -> 224x320: s = min(320/499, 224/374) = 0.598930 (height-limited), t = (int(10.57), 0) — SURVEY.md §8c."""
    scale, tr, inv = preprocess_ref.letterbox_params(people_u8.shape[:2], (224, 320))
    assert abs(scale[0] - 224 / 374) < 1e-15 and scale[0] == scale[1]
    assert tr.tolist() == [10, 0]
    assert abs(inv[0, 0] - 374 / 224) < 1e-12 and abs(inv[0, 2] + 10 * 374 / 224) < 1e-12 and inv[1, 2] == 0.0
    # the host mirror used by the product path builds the same numbers
    s2, t2, inv2 = preprocess.letterbox_params(people_u8.shape[:2], (224, 320))
    np.testing.assert_array_equal(scale, s2)
    np.testing.assert_array_equal(tr, t2)
    np.testing.assert_array_equal(inv, inv2)


def test_zero_fill_bands_and_hand_computed_pixels(people_u8):
    out = preprocess_ref.letterbox(people_u8, (224, 320))
    assert out.shape == (224, 320, 3) and out.dtype == np.uint8
    # x = (c - 10) * 1.6696...: columns 0..9 map left of the image -> fill; x(308) = 497.55 still blends source columns
    # 497 / 498, x(309) = 499.2 is past the last column -> fill from 309 on
    assert not out[:, :10].any() and not out[:, 309:].any()
    assert out[:, 10].any() and out[:, 308].any()
    # column 10 / row 0 samples the source pixel (0, 0) exactly
    np.testing.assert_array_equal(out[0, 10], people_u8[0, 0])
    # one interior pixel by hand (float64 bilinear between floor/ceil neighbours, truncation)
    r, c = 100, 150
    _, _, inv = preprocess_ref.letterbox_params(people_u8.shape[:2], (224, 320))
    x, y = inv[0, 0] * c + inv[0, 2], inv[1, 1] * r
    x0, y0, x1, y1 = int(np.floor(x)), int(np.floor(y)), int(np.ceil(x)), int(np.ceil(y))
    dc, dr = x - x0, y - y0
    src = people_u8.astype(np.float64)
    top = (1 - dc) * src[y0, x0] + dc * src[y0, x1]
    bot = (1 - dc) * src[y1, x0] + dc * src[y1, x1]
    np.testing.assert_array_equal(out[r, c], ((1 - dr) * top + dr * bot).astype(np.uint8))


def test_clip_keeps_fill_and_raises_border_blends():
    """skimage's clip step: output limited to the input's [min, max]; exact-zero fill survives when min > 0."""
    rng = np.random.default_rng(5)
    img = rng.integers(40, 200, (30, 50, 3), dtype=np.uint8)   # min >= 40 > 0
    out = preprocess_ref.letterbox(img, (64, 64))               # width-limited: bands above / below
    nz = out[out != 0]
    assert nz.min() >= img.min() and nz.max() <= img.max()
    assert (out == 0).any()
    # an image containing 0: plain clip, border blends below the smallest non-zero value may appear
    img[0, 0, 0] = 0
    out2 = preprocess_ref.letterbox(img, (64, 64))
    assert out2.max() <= img.max()


def test_process_img_is_letterbox_over_max(people_u8):
    x = preprocess_ref.process_img(people_u8, (224, 320))
    lb = preprocess_ref.letterbox(people_u8, (224, 320))
    assert x.dtype == np.float64 and x.max() == 1.0
    np.testing.assert_array_equal(x, lb / lb.max())


def test_people_golden_fixture_reproduced(people_u8, golden_weights, voc_anchors):
    """tests/golden/people_golden.json (made by tests/golden/make_golden.py --people-only): letterbox bytes, and the five
    class-14 detections of asset/people_res.jpg from the oracle network + KERAS-dialect decode on the letterboxed image."""
    import hashlib
    import json
    import os

    from oracle import decode_ref, keras_ref

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "people_golden.json")) as fh:
        gold = json.load(fh)
    lb = preprocess_ref.letterbox(people_u8, (224, 320))
    assert hashlib.sha256(lb.tobytes()).hexdigest() == gold["letterbox_sha256"]
    x = (lb / np.max(lb)).astype(np.float32)[None]
    heads = keras_ref.forward("yolo_mobilev1", golden_weights, x, alpha=0.75)
    h = decode_ref.HelperRef(voc_anchors, [224, 320], [7, 10, 14, 20], 20)
    yp = [hd[0].reshape(hd.shape[1], hd.shape[2], 3, 25) for hd in heads]
    det = decode_ref.detect_image(yp, h, [224, 320], people_u8.shape[:2], 0.7, 0.5)
    assert [(int(d[0]), int(d[1])) for d in det] == [(d[0], d[1]) for d in gold["detections"]]
    for d, g in zip(det, gold["detections"]):
        assert abs(float(d[2]) - g[2]) < 1e-4
        assert np.abs(np.array([float(v) for v in d[3:]]) - np.array(g[3:])).max() < 1e-2
    # SURVEY.md §8c secondary known answer (cv2-based there, so only loosely comparable): same five boxes, class 14
    assert [d[1] for d in gold["detections"]] == [637, 778, 745, 676, 590]
    survey = {637: (0.998, [98.7, 20.5, 316.3, 92.5]), 778: (0.987, [127.7, 206.6, 373.2, 282.0]), 745: (0.978, [139.6, 456.8, 312.8, 502.8]),
              676: (0.941, [147.7, 361.6, 273.1, 415.1]), 590: (0.802, [123.4, 128.9, 217.9, 162.5])}
    for g in gold["detections"]:
        sc, box = survey[g[1]]
        assert abs(g[2] - sc) < 0.02 and np.abs(np.array(g[3:]) - np.array(box)).max() < 0.5
