"""Inference-side mirror of ``tools.utils.Helper`` (/root/reference/tools/utils.py:53-82, 233-271, 339-406).

Holds what ``keras_inference.py:main`` takes from the helper: ``in_hw``, ``out_hw``, ``anchors`` ([L,A,2]
(w,h) as a fraction of the network input, loaded from ``data/{train_set}_anchor.npy``), ``class_num``,
``colormap``; plus ``_read_img`` / ``_process_img`` (aspect-preserving letterbox, zero fill, ``img/np.max(img)``).
The training-side members (dataset lists, augmenter, label encoding, loss helpers) are out of scope.
"""
from __future__ import annotations

import numpy as np


class Helper(object):
    def __init__(self, image_ann, class_num: int, anchors, in_hw, out_hw, validation_split=0.1):
        self.in_hw = np.array(in_hw)
        assert self.in_hw.ndim == 2
        self.out_hw = np.array(out_hw)
        assert self.out_hw.ndim == 2
        if image_ann is not None:
            raise NotImplementedError("training data lists are out of scope for the inference path")
        self.grid_wh = (1 / self.out_hw)[:, [1, 0]]
        self.class_num = class_num
        self.anchors = np.load(anchors) if isinstance(anchors, (str, bytes)) or hasattr(anchors, "__fspath__") \
            else np.asarray(anchors)
        self.anchor_number = len(self.anchors[0])
        self.output_number = len(self.anchors)
        if self.output_number != len(self.out_hw):
            raise ValueError(f"anchor file has {self.output_number} layers, output_size gives {len(self.out_hw)}")
        self.colormap = _COLORMAP

    def _read_img(self, img_path: str) -> np.ndarray:
        """tools/utils.py:339-355 (skimage.io.imread -> RGB uint8, grey -> 3 channels, alpha dropped)."""
        from PIL import Image
        img = np.array(Image.open(img_path))
        if img.ndim != 3:
            img = np.stack([img] * 3, axis=-1)
        return img[..., :3]

    def _process_img(self, img: np.ndarray, true_box=None, is_training: bool = False, is_resize: bool = True):
        """tools/utils.py:357-406, inference branch: letterbox to in_hw[0] then ``img / np.max(img)``.

        The affine parameters follow the reference exactly (scale = min(in_wh/img_wh), translation =
        ((in_wh - img_wh*scale)/2).astype(int)).  The resampling itself is skimage 0.15's ``warp``
        (order-1, constant 0 fill; third-party, not available offline); it is restated here with
        cv2.warpAffine bilinear — identical when the image already has the network size (identity warp),
        approximate otherwise (SURVEY.md §8c "secondary, approximate known answer").
        """
        if is_training:
            raise NotImplementedError("augmentation is out of scope for the inference path")
        if is_resize:
            img_wh = np.array([img.shape[1], img.shape[0]])
            in_wh = self.in_hw[0][::-1]
            scale = in_wh / img_wh
            scale[:] = np.min(scale)
            translation = ((in_wh - img_wh * scale) / 2).astype(int)
            if not (scale[0] == 1.0 and translation[0] == 0 and translation[1] == 0
                    and img.shape[0] == self.in_hw[0][0] and img.shape[1] == self.in_hw[0][1]):
                import cv2
                m = np.array([[scale[0], 0, translation[0]], [0, scale[1], translation[1]]], np.float64)
                img = cv2.warpAffine(img, m, (int(in_wh[0]), int(in_wh[1])), flags=cv2.INTER_LINEAR,
                                     borderMode=cv2.BORDER_CONSTANT, borderValue=0).astype("uint8")
        img = img / np.max(img)
        return img, true_box


_COLORMAP = [
    (255, 82, 0), (0, 255, 245), (0, 61, 255), (0, 255, 112), (0, 255, 133), (255, 0, 0), (255, 163, 0),
    (255, 102, 0), (194, 255, 0), (0, 143, 255), (51, 255, 0), (0, 82, 255), (0, 255, 41), (0, 255, 173),
    (10, 0, 255), (173, 255, 0), (0, 255, 153), (255, 92, 0), (255, 0, 255), (255, 0, 245),
] + [((37 * i) % 256, (91 * i + 60) % 256, (173 * i + 120) % 256) for i in range(60)]
