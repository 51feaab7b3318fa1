"""Drop-in for /root/reference/keras_inference.py — same CLI, same ``main(...)`` signature, same printed
result lines — with the network, the decode and the per-class NMS executed on a B200 through
libk210yolo_b200.so (hand-written sm_100a CUDA) instead of TensorFlow eager ops.

    python3 keras_inference.py PRE_CKPT TEST_IMAGE [--train_set voc --class_num 20 --model_def yolo_mobilev1
        --depth_multiplier 0.75 --image_size 224 320 --output_size 7 10 14 20 --obj_thresh 0.7 --iou_thresh 0.5]

Reference flow kept (keras_inference.py:75-176): Helper(anchors from data/{train_set}_anchor.npy) ->
network = eval(model_def) -> load_weights -> read + letterbox + normalise the image -> predict -> decode every
output layer -> score mask ``>= obj_thresh`` -> per-class NMS (max 30, IoU ``> iou_thresh``) -> print
``[top left bottom right score class]`` lines (or ``NOTE no boxes detected``) -> draw.  The per-op decode /
NMS loop of :94-131 is one fused kernel launch here (k2y_detect_keras).
"""
import argparse
import os
import sys

import numpy as np
import torch

from k210_yolo_framework_b200 import Helper, KerasDetector
from k210_yolo_framework_b200.yolonet import *  # noqa: F401,F403  (eval(model_def), as the reference does)

INFO = "[ INFO  ]"
NOTE = "[ NOTE  ]"


def detect(ckpt_weights, image_size, output_size, model_def, class_num, depth_multiplier, obj_thresh, iou_thresh,
           train_set, test_image):
    """Runs the whole path and returns (orig_img, detections) — detections as
    ``(class, flat_index, score, top, left, bottom, right)`` in the reference's output order."""
    h = Helper(None, class_num, f'data/{train_set}_anchor.npy', np.reshape(np.array(image_size), (-1, 2)),
               np.reshape(np.array(output_size), (-1, 2)))
    network = eval(model_def)  # type :yolo_mobilev2
    yolo_model, yolo_model_warpper = network([image_size[0], image_size[1], 3], len(h.anchors[0]), class_num,
                                             alpha=depth_multiplier)
    yolo_model_warpper.load_weights(str(ckpt_weights))
    print(INFO, f' Load CKPT {str(ckpt_weights)}')
    orig_img = h._read_img(str(test_image))
    image_shape = orig_img.shape[0:2]
    # reference: h._process_img(orig_img, None, False, True) on the CPU, then predict (:84-88).  Here the decoded uint8
    # image goes to the GPU as is: letterbox kernel -> uint8 front end (img / np.max(img) fused into the first conv)
    x_u8 = h.letterbox_device(orig_img)
    y_pred = yolo_model_warpper.predict_device_u8(x_u8[None])

    grid = [tuple(int(v) for v in t.shape[1:3]) for t in y_pred]
    if [tuple(int(v) for v in hw) for hw in h.out_hw] != grid:
        raise ValueError(f'--output_size {h.out_hw.tolist()} does not match the network grids {grid}')
    det = KerasDetector(h.anchors, image_size, h.out_hw, class_num, obj_thresh, iou_thresh, max_per_class=30, max_batch=1)
    dets, counts = det.run([t.contiguous() for t in y_pred], image_shape)
    return orig_img, h, KerasDetector.to_host(dets, counts)[0]


def main(ckpt_weights, image_size, output_size, model_def, class_num, depth_multiplier, obj_thresh, iou_thresh,
         train_set, test_image):
    orig_img, h, found = detect(ckpt_weights, image_size, output_size, model_def, class_num, depth_multiplier,
                                obj_thresh, iou_thresh, train_set, test_image)
    image_shape = orig_img.shape[0:2]
    if len(found) > 0:
        print(f'[top\tleft\tbottom\tright\tscore\tclass]')
        for c, _idx, score, top, left, bottom, right in found:
            print(f'[{top:.1f}\t{left:.1f}\t{bottom:.1f}\t{right:.1f}\t{score:.2f}\t{int(c):2d}]')
        if os.environ.get('K2Y_DRAW'):  # the reference always draws and calls pil_img.show() (:137-174)
            _draw(orig_img, image_shape, found, h, os.environ['K2Y_DRAW'])
    else:
        print(NOTE, ' no boxes detected')


def _draw(orig_img, image_shape, found, h, out_path):
    """keras_inference.py:137-174 — rectangles + '{class} {score:.2f}' labels (k210_yolo_framework_b200/draw.py); saved, not shown."""
    from k210_yolo_framework_b200.draw import draw_detections
    draw_detections(orig_img, found, h.colormap).save(out_path)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('--train_set', type=str, help='trian file lists', default='voc')
    parser.add_argument('--class_num', type=int, help='trian class num', default=20)
    parser.add_argument('--model_def', type=str, help='Model definition.', default='yolo_mobilev2')
    parser.add_argument('--depth_multiplier', type=float, help='mobilenet depth_multiplier', choices=[0.5, 0.75, 1.0], default=1.0)
    parser.add_argument('--image_size', type=int, help='net work input image size', default=(224, 320), nargs='+')
    parser.add_argument('--output_size', type=int, help='net work output image size', default=(7, 10, 14, 20), nargs='+')
    parser.add_argument('--obj_thresh', type=float, help='obj mask thresh', default=0.7)
    parser.add_argument('--iou_thresh', type=float, help='iou mask thresh', default=0.3)
    parser.add_argument('pre_ckpt', type=str, help='pre-train weights path')
    parser.add_argument('test_image', type=str, help='test image path')
    args = parser.parse_args(sys.argv[1:])
    main(args.pre_ckpt,
         args.image_size,
         args.output_size,
         args.model_def,
         args.class_num,
         args.depth_multiplier,
         args.obj_thresh,
         args.iou_thresh,
         args.train_set,
         args.test_image)
