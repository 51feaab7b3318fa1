"""§8f rows on the GPU: precision/recall counters (tools/custom.py) and the directory -> batches -> detections front-end."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN


def test_precision_recall_counters_match_the_reference_metric():
    from k210_yolo_framework_b200.evaluate import PrecisionRecall
    rng = np.random.default_rng(0)
    C = 20
    y_true = [rng.random((4, 7, 10, 3, 5 + C)).astype(np.float32), rng.random((4, 14, 20, 3, 5 + C)).astype(np.float32)]
    y_pred = [rng.normal(0, 2, t.shape).astype(np.float32) for t in y_true]
    for sig in (False, True):
        pr = PrecisionRecall(C, threshold=0.5, apply_sigmoid=sig)
        pr.update([torch.from_numpy(t).cuda() for t in y_true], [torch.from_numpy(t).cuda() for t in y_pred])
        tp = fp = fn = 0
        for t, p in zip(y_true, y_pred):
            tc = t[..., 4] > 0.5
            pv = 1.0 / (1.0 + np.exp(-p[..., 4].astype(np.float64))) if sig else p[..., 4]
            pc = pv > 0.5
            tp += int((tc & pc).sum())
            fp += int((~tc & pc).sum())
            fn += int((tc & ~pc).sum())
        assert [int(v) for v in pr.counts.cpu()] == [tp, fp, fn]
        prec, rec = pr.result()
        assert abs(prec - tp / (tp + fp)) < 1e-12 and abs(rec - tp / (tp + fn)) < 1e-12
    pr.reset()
    assert pr.result() == (0.0, 0.0)


def test_serving_front_end_over_a_directory(tmp_path, golden_weights, dog_u8, people_u8):
    """serve.py: five image files of two sizes -> batches of 2 -> one JSON line per image, in file order; dog.jpg's line holds
    the golden detections."""
    from PIL import Image
    import serve
    from k210_yolo_framework_b200.hdf5_write import save_keras_weights
    ckpt = tmp_path / "model.h5"
    save_keras_weights(str(ckpt), golden_weights)                   # the writer's file through the real load path
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    for i, arr in enumerate([dog_u8, people_u8, dog_u8[:, ::-1].copy(), people_u8, dog_u8]):
        Image.fromarray(arr).save(imgs / f"{i:02d}.png")
    out = tmp_path / "dets.jsonl"

    class A:
        pass
    a = A()
    a.ckpt, a.images, a.model_def, a.depth_multiplier, a.image_size, a.class_num = str(ckpt), str(imgs), "yolo_mobilev1", 0.75, (224, 320), 20
    a.anchors, a.batch, a.obj_thresh, a.iou_thresh = os.path.join(GOLDEN, "voc_anchor.npy"), 2, 0.7, 0.5
    with open(out, "w") as fh:
        n = serve.serve(a, fh)
    lines = [json.loads(l) for l in open(out)]
    assert n == 5 and [os.path.basename(l["image"]) for l in lines] == [f"{i:02d}.png" for i in range(5)]
    assert [(d["class"], d["index"]) for d in lines[0]["detections"]] == [(6, 53), (11, 765)]
    assert lines[4]["detections"] == lines[0]["detections"]
    assert [d["class"] for d in lines[1]["detections"]] == [14] * 5 and lines[3]["detections"] == lines[1]["detections"]
    assert abs(lines[0]["detections"][1]["score"] - 0.99633) < 1e-3
