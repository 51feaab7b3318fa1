"""tcgen05 tensor-core conv kernel vs an fp64 torch-CPU convolution (and vs the fp32 CUDA-core kernel),
one layer at a time through the k2y_conv2d hook, over the GEMM shapes the four networks produce."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from k210_yolo_framework_b200 import _lib
from k210_yolo_framework_b200._lib import check, lib


def _run(x0, x1, res, kernel, scale, shift, up0, ksize, stride, pad_mode, act, alpha, math):
    B = x0.shape[0]
    h, w = (x0.shape[1] * 2, x0.shape[2] * 2) if up0 else (x0.shape[1], x0.shape[2])
    c0, c1 = x0.shape[3], (x1.shape[3] if x1 is not None else 0)
    cout = kernel.shape[3]
    if pad_mode == 0:
        oh, ow = h, w
    elif pad_mode == 1:
        oh, ow = (h + 2 - ksize) // stride + 1, (w + 2 - ksize) // stride + 1
    else:
        oh, ow = (h + 1 - ksize) // stride + 1, (w + 1 - ksize) // stride + 1
    d0 = torch.from_numpy(x0).cuda()
    d1 = torch.from_numpy(x1).cuda() if x1 is not None else None
    dr = torch.from_numpy(res).cuda() if res is not None else None
    out = torch.full((B, oh, ow, cout), float("nan"), device="cuda")
    k = np.ascontiguousarray(kernel, np.float32)
    sc, sh = np.ascontiguousarray(scale, np.float32), np.ascontiguousarray(shift, np.float32)
    st = torch.cuda.current_stream()
    check(lib.k2y_conv2d(d0.data_ptr(), d1.data_ptr() if d1 is not None else None, dr.data_ptr() if dr is not None else None,
                         out.data_ptr(), k.ctypes.data, sc.ctypes.data, sh.ctypes.data, B, h, w, c0, c1, int(up0), cout, ksize,
                         stride, pad_mode, act, float(alpha), math, ctypes.c_void_p(st.cuda_stream)))
    return out.cpu().numpy()


def _ref(x0, x1, res, kernel, scale, shift, up0, ksize, stride, pad_mode, act, alpha):
    t0 = torch.from_numpy(x0).double().permute(0, 3, 1, 2)
    if up0:
        t0 = t0.repeat_interleave(2, 2).repeat_interleave(2, 3)
    x = t0 if x1 is None else torch.cat([t0, torch.from_numpy(x1).double().permute(0, 3, 1, 2)], 1)
    if pad_mode == 0:
        x = F.pad(x, (ksize // 2,) * 4)
    elif pad_mode == 1:
        x = F.pad(x, (1, 1, 1, 1))
    else:
        x = F.pad(x, (1, 0, 1, 0))
    wt = torch.from_numpy(kernel).double().permute(3, 2, 0, 1)
    y = F.conv2d(x, wt, None, stride=stride)
    y = y * torch.from_numpy(scale).double()[None, :, None, None] + torch.from_numpy(shift).double()[None, :, None, None]
    if act == 1:
        y = torch.where(y >= 0, y, y * float(np.float32(alpha)))
    elif act == 2:
        y = y.clamp(min=0)
    elif act == 3:
        y = y.clamp(0, 6)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + torch.from_numpy(res).double()
    return y.numpy()


# (B, H, W, C0, C1, up0, Cout, k, stride, pad_mode, act, residual)   H,W = stored size of src0
SHAPES = [
    (2, 28, 40, 96, 0, 0, 192, 1, 1, 0, 1, False),     # conv_pw_4 geometry: plain 1x1 via TMA, one n-tile
    (2, 56, 80, 24, 0, 0, 48, 1, 1, 0, 1, False),      # K = 24 < one k-block (TMA zero-fills), N = 48
    (3, 14, 20, 384, 0, 0, 384, 1, 1, 0, 1, False),    # conv_pw_7: several n-tiles, 12 k-blocks, ragged M (840 rows)
    (2, 7, 10, 768, 0, 0, 768, 1, 1, 0, 1, False),     # conv_pw_13: M = 140 (< 2 tiles)
    (2, 7, 10, 192, 0, 0, 75, 1, 1, 0, 0, False),      # head output: N = 75 (scalar epilogue, BN = 80)
    (2, 14, 20, 124, 0, 0, 24, 1, 1, 0, 0, True),      # mobilenet-v2 block_2 project: K = 124, N = 24, residual
    (2, 7, 10, 768, 0, 0, 192, 3, 1, 0, 1, False),     # head-1 3x3: gather path, K = 6912
    (2, 7, 10, 128, 384, 1, 128, 3, 1, 0, 1, False),   # head-2 3x3: upsample(128) || 384 concat
    (2, 16, 16, 64, 0, 0, 128, 3, 2, 2, 1, False),     # darknet downsample: pad((1,0),(1,0)) stride 2
    (2, 12, 12, 64, 0, 0, 64, 3, 1, 0, 1, True),       # darknet resblock 3x3 with residual add
    (2, 6, 6, 256, 512, 1, 256, 1, 1, 0, 1, False),    # make_last_layers first 1x1 on upsample || skip concat (gather, k=1)
    (1, 13, 13, 1024, 0, 0, 256, 1, 1, 0, 1, False),   # tiny_yolo 1x1 1024->256
    (2, 26, 30, 16, 0, 0, 32, 3, 1, 0, 1, False),      # tiny_yolo 3x3 16->32: two 16-channel taps per k-block, K = 144 padded to 160
    (1, 17, 19, 16, 0, 0, 48, 3, 2, 2, 1, False),      # same gather form, stride 2 with pad((1,0),(1,0)), odd extents
    (2, 19, 21, 32, 0, 0, 64, 3, 1, 0, 1, True),       # darknet 3x3 32->64 resblock: two 32-channel taps per bf16 k-block, K = 288 -> 320
    (1, 24, 20, 32, 0, 0, 64, 3, 2, 2, 1, False),      # darknet 3x3 32->64 stride 2
]


@pytest.mark.parametrize("math", [_lib.MATH_TC_3XTF32, _lib.MATH_TC_TF32, _lib.MATH_TC_BF16X3])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"s{i}" for i in range(len(SHAPES))])
def test_tc_conv_matches_reference(shape, math):
    B, H, W, C0, C1, up0, Cout, k, stride, pad_mode, act, use_res = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    x0 = rng.normal(0, 1, (B, H, W, C0)).astype(np.float32)
    hh, ww = (H * 2, W * 2) if up0 else (H, W)
    x1 = rng.normal(0, 1, (B, hh, ww, C1)).astype(np.float32) if C1 else None
    kernel = (rng.normal(0, 1, (k, k, C0 + C1, Cout)) / np.sqrt(k * k * (C0 + C1))).astype(np.float32)
    scale = rng.uniform(0.5, 2.0, Cout).astype(np.float32)
    shift = rng.normal(0, 0.5, Cout).astype(np.float32)
    ref0 = _ref(x0, x1, None, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1)
    res = rng.normal(0, 1, ref0.shape).astype(np.float32) if use_res else None
    ref = _ref(x0, x1, res, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1)
    got = _run(x0, x1, res, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1, math)
    assert got.shape == ref.shape
    assert np.isfinite(got).all(), "output has unwritten (NaN) elements"
    err = np.abs(got - ref).max()
    # The tensor core rounds its fp32 accumulator toward zero once per tcgen05.mma (K = 8): a chain of
    # 3*K/8 (3xTF32) instructions loses up to ~2^-24 * chain * |acc|.  Measured: 2.4e-5 at K=384, 3.5e-4 at K=6912.
    chain = 3 * (k * k * (C0 + C1)) / 8
    if math == _lib.MATH_TC_3XTF32:
        tol = max(5e-5, 2.0 ** -24 * chain * float(np.abs(ref).max()))
    elif math == _lib.MATH_TC_BF16X3:
        tol = 6e-4   # (hi, mid) bf16 planes: ~2^-16 relative per product, random-walk over K
    else:
        tol = 2e-2
    assert err < tol, f"max abs err {err:.3e} (tol {tol:.3e})"
    simt = _run(x0, x1, res, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1, _lib.MATH_FP32_SIMT)
    assert np.abs(simt - ref).max() < 2e-5


def test_unsupported_shape_is_reported():
    x0 = np.zeros((1, 8, 8, 3), np.float32)
    kernel = np.zeros((3, 3, 3, 16), np.float32)
    with pytest.raises(_lib.K2YError):
        _run(x0, None, None, kernel, np.ones(16), np.zeros(16), 0, 3, 1, 0, 0, 0.0, _lib.MATH_TC_3XTF32)


@pytest.mark.parametrize("splits", [2, 4, 7])
def test_split_k_matches_reference(splits, monkeypatch):
    """Deep-K 3x3 convs split their K range over several CTAs (partials through the TMA-store epilogue + reduce pass)."""
    monkeypatch.setenv("K2Y_TC_SPLITK", str(splits))
    for shape in (SHAPES[6], SHAPES[7], SHAPES[9]):
        B, H, W, C0, C1, up0, Cout, k, stride, pad_mode, act, use_res = shape
        rng = np.random.default_rng(splits)
        x0 = rng.normal(0, 1, (B, H, W, C0)).astype(np.float32)
        hh, ww = (H * 2, W * 2) if up0 else (H, W)
        x1 = rng.normal(0, 1, (B, hh, ww, C1)).astype(np.float32) if C1 else None
        kernel = (rng.normal(0, 1, (k, k, C0 + C1, Cout)) / np.sqrt(k * k * (C0 + C1))).astype(np.float32)
        scale = rng.uniform(0.5, 2.0, Cout).astype(np.float32)
        shift = rng.normal(0, 0.5, Cout).astype(np.float32)
        ref0 = _ref(x0, x1, None, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1)
        res = rng.normal(0, 1, ref0.shape).astype(np.float32) if use_res else None
        ref = _ref(x0, x1, res, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1)
        got = _run(x0, x1, res, kernel, scale, shift, up0, k, stride, pad_mode, act, 0.1, _lib.MATH_TC_3XTF32)
        assert np.isfinite(got).all()
        assert np.abs(got - ref).max() < 2e-4, f"splits={splits} shape={shape}: {np.abs(got - ref).max():.3e}"
