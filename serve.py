#!/usr/bin/env python
"""Batched serving front-end: a directory (or list) of images -> batches -> image-shard over the GPUs of the box ->
gathered detections, one JSON line per image.  The GPU analogue of the firmware's capture -> run -> draw loop
(/root/reference/yolo3_frame_test_public/main.c:294-326) and of calling keras_inference.py once per file.

    python serve.py CKPT IMAGE_DIR [--model_def yolo_mobilev1 --depth_multiplier 0.75 --image_size 224 320 --class_num 20
                                    --anchors data/voc_anchor.npy --batch 32 --obj_thresh 0.7 --iou_thresh 0.5 --out dets.jsonl]
    python -m torch.distributed.run --nproc-per-node 8 serve.py ...        # one rank per GPU, NCCL all-gather of the records

Per global batch of world*batch images: every rank decodes ITS images on the host (PIL), letterboxes each on the GPU
(k2y_letterbox_u8) into its uint8 input buffer, runs network + decode + NMS, and ONE all-gather completes the record blocks on
every rank; rank 0 writes the lines in file order.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from k210_yolo_framework_b200 import Helper  # noqa: E402
from k210_yolo_framework_b200.dist import shard_range  # noqa: E402
from k210_yolo_framework_b200.pipeline import DetectionPipeline  # noqa: E402

EXTS = (".jpg", ".jpeg", ".png", ".bmp")


def list_images(path):
    if os.path.isdir(path):
        return sorted(os.path.join(path, f) for f in os.listdir(path) if f.lower().endswith(EXTS))
    with open(path) as fh:
        return [l.strip() for l in fh if l.strip()]


def serve(args, out_fh=None):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    anchors = np.load(args.anchors)
    H, W = args.image_size
    out_hw = [(H // 32 * 2 ** l, W // 32 * 2 ** l) for l in range(len(anchors))]
    h = Helper(None, args.class_num, anchors, np.reshape(np.array(args.image_size), (-1, 2)), np.reshape(np.array(out_hw), (-1, 2)))
    pipe = DetectionPipeline(args.model_def, (H, W), anchors, args.class_num, args.depth_multiplier, args.batch, args.obj_thresh,
                             args.iou_thresh, 30, device=local, world=world, rank=rank)
    pipe.engine.load_weights(args.ckpt)
    files = list_images(args.images)
    buf = torch.zeros((args.batch, H, W, 3), dtype=torch.uint8, device=f"cuda:{local}")
    shapes = np.tile(np.array([[H, W]], np.float32), (args.batch, 1))
    G = world * args.batch
    n_out = 0
    for g0 in range(0, len(files), G):
        chunk = files[g0:g0 + G]
        lo, hi = rank * args.batch, min((rank + 1) * args.batch, len(chunk))
        for j in range(lo, hi):                                   # this rank's images of the global batch
            img = h._read_img(chunk[j])
            shapes[j - lo] = img.shape[:2]
            buf[j - lo].copy_(h.letterbox_device(img))
        pipe.set_image_shapes(shapes)
        pipe.engine.bind_input(buf)
        dets, counts = pipe.step_device(args.batch)
        pipe.wait_gathered()
        torch.cuda.synchronize()
        if rank == 0:
            recs = DetectionPipeline.records(dets[:len(chunk)].clone() if world == 1 else dets.clone(), counts.clone())
            for j, path in enumerate(chunk):
                line = {"image": path, "detections": [{"class": int(c), "score": round(float(s), 6), "box_tlbr": [float(t), float(l), float(b), float(r)],
                                                       "index": int(i)} for c, i, s, t, l, b, r in recs[j]]}
                if out_fh is not None:
                    out_fh.write(json.dumps(line) + "\n")
                n_out += 1
    if world > 1:
        dist.barrier()
    return n_out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ckpt")
    ap.add_argument("images", help="directory of images, or a text file with one path per line")
    ap.add_argument("--model_def", default="yolo_mobilev1", choices=["yolo_mobilev1", "yolo_mobilev2", "tiny_yolo", "yolo"])
    ap.add_argument("--depth_multiplier", type=float, default=0.75)
    ap.add_argument("--image_size", type=int, nargs=2, default=(224, 320))
    ap.add_argument("--class_num", type=int, default=20)
    ap.add_argument("--anchors", default="data/voc_anchor.npy")
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--obj_thresh", type=float, default=0.7)
    ap.add_argument("--iou_thresh", type=float, default=0.5)
    ap.add_argument("--out", default="-")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    fh = (sys.stdout if args.out == "-" else open(args.out, "w")) if rank == 0 else None
    n = serve(args, fh)
    if rank == 0 and fh is not sys.stdout:
        fh.close()
        print(f"{n} images -> {args.out}", file=sys.stderr)


if __name__ == "__main__":
    main()
