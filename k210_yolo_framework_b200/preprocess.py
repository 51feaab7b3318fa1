"""GPU pre-processing mirror of ``Helper._process_img`` (tools/utils.py:357-406, inference branch).

``letterbox_params`` builds the affine parameters exactly as the reference does (numpy, a handful of host flops);
``letterbox_device`` runs the resampling on the GPU through the C-ABI (``k2y_letterbox_u8``, csrc/preprocess.cu) and
returns the uint8 network input; the ``img / np.max(img)`` normalisation is fused into the network's uint8 front end
(``YoloEngine.predict_device_u8``).  There is no CPU fallback: without the CUDA library the call raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from ._lib import check, lib


def letterbox_params(img_hw, in_hw):
    """tools/utils.py:374-381,399: scale = min(in_wh / img_wh) on both axes, translation = ((in_wh - img_wh*scale)/2).astype(int),
    and inv(AffineTransform(scale, translation).params) — the matrix ``aff.inverse`` gives to ``skimage.transform.warp``."""
    img_wh = np.array([img_hw[1], img_hw[0]])
    in_wh = np.array([in_hw[1], in_hw[0]])
    scale = in_wh / img_wh  # NOTE affine transform scale is [w, h]
    scale[:] = np.min(scale)
    translation = ((in_wh - img_wh * scale) / 2).astype(int)
    fwd = np.array([[scale[0], 0.0, float(translation[0])], [0.0, scale[1], float(translation[1])], [0.0, 0.0, 1.0]])
    return scale, translation, np.linalg.inv(fwd)


def letterbox_device(img_u8: torch.Tensor, in_hw, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """img_u8: CUDA uint8 [h, w, 3] RGB of any size -> CUDA uint8 [in_h, in_w, 3] letterboxed (zero fill)."""
    if img_u8.dim() != 3 or img_u8.shape[2] != 3 or img_u8.dtype != torch.uint8 or not img_u8.is_cuda:
        raise ValueError(f"expected CUDA uint8 [h,w,3], got {tuple(img_u8.shape)} {img_u8.dtype} {img_u8.device}")
    img_u8 = img_u8.contiguous()
    in_h, in_w = int(in_hw[0]), int(in_hw[1])
    if out is None:
        out = torch.empty((in_h, in_w, 3), dtype=torch.uint8, device=img_u8.device)
    elif tuple(out.shape) != (in_h, in_w, 3) or out.dtype != torch.uint8 or out.device != img_u8.device or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA uint8 [in_h,in_w,3] tensor on the image's device")
    _, _, inv = letterbox_params(img_u8.shape[:2], (in_h, in_w))
    m = (ctypes.c_double * 6)(*[float(v) for v in inv[:2].reshape(-1)])
    minmax = torch.empty((2,), dtype=torch.int32, device=img_u8.device)
    with torch.cuda.device(img_u8.device):
        st = torch.cuda.current_stream()
        check(lib.k2y_letterbox_u8(img_u8.data_ptr(), int(img_u8.shape[0]), int(img_u8.shape[1]), m, out.data_ptr(), in_h, in_w,
                                   minmax.data_ptr(), ctypes.c_void_p(st.cuda_stream)))
    return out
