"""Generates the committed golden fixtures from the reference's shipped assets.

Run once in the build container (``python tests/golden/make_golden.py``); needs
/root/reference, which does not exist on the GPU box — hence the committed outputs.

Inputs (reference fixtures):
  asset/yolo_model.h5   trained yolo_mobilev1 alpha=0.75, VOC-20   -> weights npz (same f32 values)
  data/dog.jpg, data/people.jpg                                    -> decoded uint8 RGB arrays
  data/voc_anchor.npy                                              -> copied values
Outputs of the oracle / compiled reference on them:
  dog_heads_f32.npz     oracle fp32 head tensors (TF-CPU stand-in) and fp64 ground truth
  dog_golden.json       KERAS-dialect detections (obj 0.7, iou 0.5) + REGION_C draw list from the
                        *compiled reference* region_layer.c (thr 0.6, nms 0.3 — main.c:280-287)
  people_golden.json    data/people.jpg (374x499) through the restated letterbox (oracle/preprocess_ref.py; skimage is
                        not available offline -> "parity unpinned" for the resampling) + oracle network + KERAS decode:
                        sha256 / sample pixels of the letterboxed uint8 input and the five class-14 detections that
                        asset/people_res.jpg shows.  `--people-only` regenerates it from the committed fixtures alone.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from k210_yolo_framework_b200.hdf5_min import load_keras_weights  # noqa: E402
from oracle import decode_ref, keras_ref, region_c  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    w = load_keras_weights(f"{REF}/asset/yolo_model.h5")
    flat = {f"{layer}/{var}": arr for layer, vs in w.items() for var, arr in vs.items()}
    np.savez_compressed(f"{OUT}/yolo_mobilev1_075_voc_weights.npz", **flat)
    anchors = np.load(f"{REF}/data/voc_anchor.npy")
    np.save(f"{OUT}/voc_anchor.npy", anchors)
    for name in ("dog", "people"):
        img = np.array(Image.open(f"{REF}/data/{name}.jpg").convert("RGB"))
        np.save(f"{OUT}/{name}_u8.npy", img)
    dog = np.load(f"{OUT}/dog_u8.npy")
    x = (dog / np.max(dog)).astype(np.float32)[None]
    heads32 = keras_ref.forward("yolo_mobilev1", w, x, alpha=0.75)
    heads64 = keras_ref.forward("yolo_mobilev1", w, x.astype(np.float64), alpha=0.75, dtype=torch.float64)
    np.savez_compressed(f"{OUT}/dog_heads.npz", l0_f32=heads32[0], l1_f32=heads32[1],
                        l0_f64=heads64[0], l1_f64=heads64[1])
    h = decode_ref.HelperRef(anchors, [224, 320], [7, 10, 14, 20], 20)
    yp = [hd[0].reshape(hd.shape[1], hd.shape[2], 3, 25) for hd in heads32]
    det = decode_ref.detect_image(yp, h, [224, 320], dog.shape[:2], 0.7, 0.5)
    gold = {"keras": {"obj_thresh": 0.7, "iou_thresh": 0.5,
                      "detections": [[int(d[0]), int(d[1])] + [float(v) for v in d[2:]] for d in det]},
            "region_c": {"threshold": 0.6, "nms_value": 0.3, "layers": []}}
    for l, (W, H) in enumerate([(10, 7), (20, 14)]):
        chw = region_c.nhwc_to_chw(heads32[l][0], 3)
        r = region_c.RegionLayerRef(W, H, 75, 320, 224, anchors[l].reshape(-1), 0.6, 0.3)
        gold["region_c"]["layers"].append([[int(v) for v in t[:5]] + [float(t[5])] for t in r.run(chw)])
    with open(f"{OUT}/dog_golden.json", "w") as fh:
        json.dump(gold, fh, indent=1)
    print(json.dumps(gold, indent=1))


def people_golden():
    from k210_yolo_framework_b200.yolonet import load_npz_weights
    from oracle import preprocess_ref
    w = load_npz_weights(f"{OUT}/yolo_mobilev1_075_voc_weights.npz")
    anchors = np.load(f"{OUT}/voc_anchor.npy")
    img = np.load(f"{OUT}/people_u8.npy")
    scale, tr, inv = preprocess_ref.letterbox_params(img.shape[:2], (224, 320))
    lb = preprocess_ref.letterbox(img, (224, 320))
    x = (lb / np.max(lb)).astype(np.float32)[None]
    heads32 = keras_ref.forward("yolo_mobilev1", w, x, alpha=0.75)
    h = decode_ref.HelperRef(anchors, [224, 320], [7, 10, 14, 20], 20)
    yp = [hd[0].reshape(hd.shape[1], hd.shape[2], 3, 25) for hd in heads32]
    det = decode_ref.detect_image(yp, h, [224, 320], img.shape[:2], 0.7, 0.5)
    samples = [[r, c] + [int(v) for v in lb[r, c]] for r, c in [(0, 10), (100, 150), (223, 309), (57, 200), (180, 11)]]
    gold = {"image_hw": list(img.shape[:2]), "in_hw": [224, 320], "scale": float(scale[0]), "translation": [int(t) for t in tr],
            "letterbox_sha256": hashlib.sha256(lb.tobytes()).hexdigest(), "letterbox_samples_r_c_rgb": samples,
            "obj_thresh": 0.7, "iou_thresh": 0.5,
            "detections": [[int(d[0]), int(d[1])] + [float(v) for v in d[2:]] for d in det]}
    with open(f"{OUT}/people_golden.json", "w") as fh:
        json.dump(gold, fh, indent=1)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    if "--people-only" not in sys.argv:
        main()
    people_golden()


def label_boxes_of_reference_render():
    """Geometry of the filled label boxes in the reference's own rendered result (asset/dog_res.jpg, drawn by
    keras_inference.py:137-174): per class, the bounding rows / columns of the pixels close to the class colour.  The 1-px
    rectangle outlines do not survive the JPEG compression, the filled label boxes do.  -> tests/golden/dog_res_labels.json"""
    import json
    from PIL import Image
    from k210_yolo_framework_b200.helper import _COLORMAP
    im = np.array(Image.open("/root/reference/asset/dog_res.jpg").convert("RGB")).astype(int)
    out = {}
    for c in (6, 11):
        d = np.abs(im - np.array(_COLORMAP[c])).sum(-1)
        ys, xs = np.nonzero(d < 60)
        out[str(c)] = {"ymin": int(ys.min()), "ymax": int(ys.max()), "xmin": int(xs.min()), "xmax": int(xs.max()), "pixels": int(len(ys))}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dog_res_labels.json"), "w") as fh:
        json.dump(out, fh)
    return out
