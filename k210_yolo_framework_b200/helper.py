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

        The resampling runs on the GPU (``preprocess.letterbox_device`` -> k2y_letterbox_u8): the affine parameters follow the
        reference exactly (scale = min(in_wh/img_wh), translation = ((in_wh - img_wh*scale)/2).astype(int)) and the
        interpolation restates skimage 0.15's ``warp`` (order 1, zero fill, clip, uint8 truncation; third-party, not
        available offline — oracle/preprocess_ref.py).  Exact for an image that already has the network size.
        Returns the reference's value: a float64 HxWx3 array in [0, 1] on the host.
        """
        if is_training:
            raise NotImplementedError("augmentation is out of scope for the inference path")
        if is_resize:
            img = self.letterbox_device(img).cpu().numpy()
        img = img / np.max(img)
        return img, true_box

    def letterbox_device(self, img: np.ndarray):
        """uint8 HWC host image -> CUDA uint8 [in_h, in_w, 3]: the input of the network's uint8 front end
        (``predict_device_u8``), which applies ``img / np.max(img)`` on the GPU."""
        import torch
        from .preprocess import letterbox_device
        x = torch.from_numpy(np.ascontiguousarray(img[..., :3], dtype=np.uint8)).cuda()
        return letterbox_device(x, self.in_hw[0])


_COLORMAP = [
    (255, 82, 0), (0, 255, 245), (0, 61, 255), (0, 255, 112), (0, 255, 133), (255, 0, 0), (255, 163, 0),
    (255, 102, 0), (194, 255, 0), (0, 143, 255), (51, 255, 0), (0, 82, 255), (0, 255, 41), (0, 255, 173),
    (10, 0, 255), (173, 255, 0), (0, 255, 153), (255, 92, 0), (255, 0, 255), (255, 0, 245),
] + [((37 * i) % 256, (91 * i + 60) % 256, (173 * i + 120) % 256) for i in range(60)]
