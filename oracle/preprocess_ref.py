"""TEST INFRASTRUCTURE — CPU restatement of the reference's inference-time image pre-processing.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path is
k210_yolo_framework_b200/csrc/preprocess.cu.

Reference: tools/utils.py:357-406 (`Helper._process_img`, inference branch):

    scale        = min(in_wh / img_wh)                       (same factor on both axes)
    translation  = ((in_wh - img_wh * scale) / 2).astype(int)
    aff          = skimage.transform.AffineTransform(scale=scale, translation=translation)
    img          = skimage.transform.warp(img, aff.inverse, output_shape=in_hw, preserve_range=True).astype('uint8')
    img          = img / np.max(img)

`skimage.transform.warp` is third-party (scikit_image==0.15.0, requirements.txt:7) and NOT available offline, so this is
"parity unpinned" for the resampling step: the algorithm below restates skimage 0.15's published behaviour for this call
(order=1, mode='constant', cval=0, clip=True, preserve_range=True, homography fast path):

  * matrix = np.linalg.inv(aff.params); for output pixel (row r, col c):  x = M00*c + M01*r + M02,  y = M10*c + M11*r + M12
    (no half-pixel centre shift), in float64;
  * bilinear: minr = floor(y), maxr = ceil(y), minc = floor(x), maxc = ceil(x), dr = y - minr, dc = x - minc,
    top = (1-dc)*P(minr,minc) + dc*P(minr,maxc), bottom = (1-dc)*P(maxr,minc) + dc*P(maxr,maxc),
    out = (1-dr)*top + dr*bottom, with P = 0 (cval) outside the image;
  * clip to [min(img), max(img)] of the whole input; when 0 lies outside that range the exactly-zero (fill) pixels stay 0;
  * `.astype('uint8')` truncates toward zero.

The identity case (image already in_hw, scale 1, translation 0 — data/dog.jpg) is exact by construction, which is what
the pinned known answers of tests/golden/dog_golden.json rely on.
"""
from __future__ import annotations

import numpy as np


def letterbox_params(img_hw, in_hw):
    """scale (float64), translation (int x, int y) and the inverse 3x3 matrix, exactly as tools/utils.py:377-400 builds them."""
    img_wh = np.array([img_hw[1], img_hw[0]])
    in_wh = np.array([in_hw[1], in_hw[0]])
    scale = in_wh / img_wh
    scale[:] = np.min(scale)
    translation = ((in_wh - img_wh * scale) / 2).astype(int)
    fwd = np.array([[scale[0], 0.0, float(translation[0])], [0.0, scale[1], float(translation[1])], [0.0, 0.0, 1.0]])
    return scale, translation, np.linalg.inv(fwd)


def warp_order1(img_u8: np.ndarray, inv: np.ndarray, out_hw) -> np.ndarray:
    """skimage-0.15-style `warp(img, inverse, output_shape, order=1, mode='constant', preserve_range=True).astype('uint8')`."""
    img = np.asarray(img_u8)
    H, W = img.shape[:2]
    oh, ow = int(out_hw[0]), int(out_hw[1])
    c = np.arange(ow, dtype=np.float64)[None, :]
    r = np.arange(oh, dtype=np.float64)[:, None]
    x = (inv[0, 0] * c + inv[0, 1] * r) + inv[0, 2]
    y = (inv[1, 0] * c + inv[1, 1] * r) + inv[1, 2]
    minr, minc = np.floor(y), np.floor(x)
    maxr, maxc = np.ceil(y), np.ceil(x)
    dr, dc = y - minr, x - minc
    src = img.astype(np.float64)

    def pix(rr, cc):
        ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
        ri = np.clip(rr, 0, H - 1).astype(np.int64)
        ci = np.clip(cc, 0, W - 1).astype(np.int64)
        v = src[ri, ci]
        return np.where(ok[..., None], v, 0.0)

    dc3, dr3 = dc[..., None], dr[..., None]
    top = (1.0 - dc3) * pix(minr, minc) + dc3 * pix(minr, maxc)
    bottom = (1.0 - dc3) * pix(maxr, minc) + dc3 * pix(maxr, maxc)
    out = (1.0 - dr3) * top + dr3 * bottom
    lo, hi = float(img.min()), float(img.max())
    if not (lo <= 0.0 <= hi):
        fill = out == 0.0
        out = np.clip(out, lo, hi)
        out[fill] = 0.0
    else:
        out = np.clip(out, lo, hi)
    return out.astype(np.uint8)


def letterbox(img_u8: np.ndarray, in_hw) -> np.ndarray:
    """uint8 HWC image of any size -> uint8 [in_h, in_w, 3] letterboxed network input (before `img / np.max(img)`)."""
    _, _, inv = letterbox_params(img_u8.shape[:2], in_hw)
    return warp_order1(img_u8, inv, in_hw)


def process_img(img_u8: np.ndarray, in_hw) -> np.ndarray:
    """tools/utils.py:357-406 inference branch end to end: float64 image in [0, 1]."""
    img = letterbox(img_u8, in_hw)
    return img / np.max(img)
