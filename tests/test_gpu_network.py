"""CUDA network path vs the oracle — through the reference-facing builder API / C-ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from k210_yolo_framework_b200 import _lib, yolonet
from k210_yolo_framework_b200.weights import random_weights
from oracle import decode_ref, keras_ref

MODES = [_lib.MATH_FP32_SIMT, _lib.MATH_TC_3XTF32, _lib.MATH_TC_BF16X3]
# per-layer error budget relative to the layer's max |activation| (fp64 oracle): fp32-class modes vs the ~16-bit-mantissa bf16x3
LAYER_TOL = {_lib.MATH_FP32_SIMT: 2e-4, _lib.MATH_TC_3XTF32: 2e-4, _lib.MATH_TC_BF16X3: 6e-4}


def _maxerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


@pytest.mark.parametrize("math", MODES)
def test_mobilev1_real_weights_layerwise(golden_weights, dog_u8, dog_heads, math):
    """Every conv layer of the trained yolo_mobilev1-0.75 on data/dog.jpg against the fp64 oracle."""
    x = (dog_u8 / np.max(dog_u8)).astype(np.float32)[None]
    m, w = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=2)
    m.engine.set_keep_all(True)
    m.engine.set_use_graph(False)
    m.set_weights_dict(golden_weights)
    m.engine.set_math(math)
    heads = m.predict(x)
    _, acts = keras_ref.forward("yolo_mobilev1", golden_weights, x.astype(np.float64), alpha=0.75,
                                dtype=torch.float64, record=True)
    worst = {}
    for L in m.engine.layers():
        name, bn = L.name.decode(), L.bn_name.decode()
        got = m.engine.read_layer(name, 1)
        # the engine stores the post-BN/activation tensor; pick the oracle's matching record
        if name.startswith("conv_dw_"):
            key = name + "_relu"
        elif name.startswith("conv_pw_") or name == "conv1":
            key = name + "_relu"
        elif bn:
            key = bn + "/leaky"
        else:
            key = name
        ref = acts[key].permute(0, 2, 3, 1).numpy()
        assert got.shape == ref.shape, name
        scale = max(1.0, float(np.abs(ref).max()))
        err = _maxerr(got, ref) / scale
        worst[name] = err
        assert err < LAYER_TOL[math], f"{name}: rel-to-max error {err:.3e} ({_lib.MATH_NAMES[math]})"
    # heads: absolute logit error well inside the 1e-3 score/box budget
    e0, e1 = _maxerr(heads[0], dog_heads["l0_f64"]), _maxerr(heads[1], dog_heads["l1_f64"])
    print(f"[{_lib.MATH_NAMES[math]}] head logit max abs error vs fp64 oracle: {e0:.2e} {e1:.2e}; worst layer {max(worst.values()):.2e}")
    assert e0 < 1e-3 and e1 < 1e-3
    # wrapper view
    hw = w.predict(x)
    assert hw[0].shape == (1, 7, 10, 3, 25) and hw[1].shape == (1, 14, 20, 3, 25)
    np.testing.assert_array_equal(hw[0].reshape(heads[0].shape), heads[0])


CASES = [("yolo_mobilev1", 0.75, 20, (224, 320), 3), ("yolo_mobilev1", 1.0, 20, (96, 128), 2),
         ("yolo_mobilev2", 1.0, 20, (224, 320), 2), ("yolo_mobilev2", 0.5, 20, (96, 96), 2),
         ("tiny_yolo", 1.0, 20, (160, 160), 2), ("yolo", 1.0, 80, (96, 96), 2)]


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("model_def,alpha,classes,hw,batch", CASES)
def test_random_weights_all_models(model_def, alpha, classes, hw, batch, math):
    m, _ = getattr(yolonet, model_def)([hw[0], hw[1], 3], 3, classes, alpha=alpha, max_batch=batch)
    weights = random_weights(m.engine.expected_variables(), seed=7, detection_rich=True)
    m.set_weights_dict(weights)
    m.engine.set_math(math)
    x = np.random.default_rng(11).random((batch, hw[0], hw[1], 3), dtype=np.float32)
    got = m.predict(x)
    ref = keras_ref.forward(model_def, weights, x.astype(np.float64), alpha=alpha, dtype=torch.float64)
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        scale = max(1.0, float(np.abs(r).max()))
        assert _maxerr(g, r) / scale < LAYER_TOL[math], f"{model_def} {_lib.MATH_NAMES[math]}: {_maxerr(g, r):.3e} (max |ref| {scale:.2f})"
    # the contract (north_star): decoded confidences and normalised box coordinates within 1e-3 of the reference's
    L = len(got)
    anchors = np.stack([[[0.76, 0.57], [0.69, 0.89], [0.47, 0.34]]] * L) / np.array([1.0, 2.0, 4.0])[:L, None, None]
    h = decode_ref.HelperRef(anchors, list(hw), [(t.shape[1], t.shape[2]) for t in got], classes)
    for b in range(batch):
        gb, gs = decode_ref.decode_layers([t[b].reshape(t.shape[1], t.shape[2], 3, 5 + classes) for t in got], h, list(hw), hw)
        rb, rs = decode_ref.decode_layers([t[b].astype(np.float32).reshape(t.shape[1], t.shape[2], 3, 5 + classes) for t in ref], h, list(hw), hw)
        assert _maxerr(gs, rs) < 1e-3, f"{model_def} {_lib.MATH_NAMES[math]}: score error {_maxerr(gs, rs):.2e}"
        norm = np.array([hw[0], hw[1], hw[0], hw[1]], np.float64)
        box_err = np.abs(gb / norm - rb / norm) / np.maximum(1.0, np.abs(rb / norm))     # relative for boxes wider than the image
        assert float(box_err.max()) < 1e-3, f"{model_def} {_lib.MATH_NAMES[math]}: normalised box error {float(box_err.max()):.2e}"


def test_device_api_graph_replay_and_batching():
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=4)
    weights = random_weights(m.engine.expected_variables(), seed=5, detection_rich=True)
    m.set_weights_dict(weights)
    x = torch.rand((4, 224, 320, 3), device="cuda")
    a = [t.clone() for t in m.predict_device(x)]          # captures the graph for batch 4
    b = [t.clone() for t in m.predict_device(x)]          # replays it
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    c = [t.clone() for t in m.predict_device(x[1:3])]     # different batch -> second graph; tile / split-K shapes may
    for p, q in zip(a, c):                                 # differ with M, so only the accumulation order changes
        assert torch.allclose(p[1:3], q, atol=2e-4, rtol=1e-4)
    m.engine.set_use_graph(False)
    d = m.predict_device(x)
    for p, q in zip(a, d):
        assert torch.equal(p, q)
    host = m.predict(x.cpu().numpy())
    for p, q in zip(a, host):
        np.testing.assert_array_equal(p.cpu().numpy(), q)
    assert m.engine.launches_per_run() >= 32   # + split-K reducers


def test_fused_depthwise_pointwise_blocks(monkeypatch):
    """Default schedule: every stride-1 MobileNet block (depthwise 3x3 + 1x1) is ONE launch of dwpw_tc_kernel — the depthwise
    result goes from a TMA-staged shared-memory window straight into tensor memory.  Same heads as the unfused schedule
    (K2Y_NO_DWPW=1), fewer launches, and the oracle on extents where tiles are partial."""
    monkeypatch.setenv("K2Y_DWPW", "1")
    monkeypatch.delenv("K2Y_NO_DWPW", raising=False)
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=3)
    weights = random_weights(m.engine.expected_variables(), seed=9, detection_rich=True)
    m.set_weights_dict(weights)
    m.engine.set_use_graph(False)
    x = torch.rand((3, 224, 320, 3), device="cuda")
    fused = [t.clone() for t in m.predict_device(x)]
    n_fused = m.engine.launches_per_run()
    names = [p["name"] for p in m.engine.profile(3)]
    for blk in (1, 3, 5, 7, 8, 9, 10, 11):
        assert f"conv_dw_{blk}+conv_pw_{blk}" in names, names
    monkeypatch.setenv("K2Y_NO_DWPW", "1")
    monkeypatch.setenv("K2Y_DWPW", "0")
    plain = [t.clone() for t in m.predict_device(x)]
    assert m.engine.launches_per_run() == n_fused + 8
    assert not any("+" in p["name"] for p in m.engine.profile(3))
    monkeypatch.delenv("K2Y_NO_DWPW")
    monkeypatch.setenv("K2Y_DWPW", "1")
    for p, q in zip(fused, plain):
        assert torch.allclose(p, q, atol=3e-4, rtol=1e-4), float((p - q).abs().max())
    # one block at a time (K2Y_DWPW_MASK selects depthwise layers by schedule order): localises a failure to a layer shape
    for blk in (1, 3, 5, 7):
        monkeypatch.setenv("K2Y_DWPW_MASK", hex(1 << (blk - 1)))
        one = m.predict_device(x)
        assert m.engine.launches_per_run() == n_fused + 7
        for p, q in zip(one, plain):
            assert torch.allclose(p, q, atol=3e-4, rtol=1e-4), (blk, float((p - q).abs().max()))
    monkeypatch.delenv("K2Y_DWPW_MASK")
    # the default schedule fuses the HBM-bound early blocks only (C <= 96), where the single launch measured faster
    monkeypatch.delenv("K2Y_DWPW")
    m.predict_device(x)
    dflt = [p["name"] for p in m.engine.profile(3) if "+" in p["name"]]
    assert dflt == ["conv_dw_1+conv_pw_1", "conv_dw_3+conv_pw_3"], dflt
    monkeypatch.setenv("K2Y_DWPW", "1")
    # against the oracle, on extents where tiles are partial (96x160 -> 48x80, 24x40, 12x20, 6x10 maps) and C = 16..512
    for alpha, hw in ((0.5, (96, 160)), (1.0, (64, 96))):
        m2, _ = yolonet.yolo_mobilev1([hw[0], hw[1], 3], 3, 20, alpha=alpha, max_batch=2)
        w2 = random_weights(m2.engine.expected_variables(), seed=10, detection_rich=True)
        m2.set_weights_dict(w2)
        x2 = np.random.default_rng(3).random((2, hw[0], hw[1], 3), dtype=np.float32)
        got = m2.predict(x2)
        assert any("+" in p["name"] for p in m2.engine.profile(2))
        ref = keras_ref.forward("yolo_mobilev1", w2, x2.astype(np.float64), alpha=alpha, dtype=torch.float64)
        for g, r in zip(got, ref):
            assert _maxerr(g, r) / max(1.0, float(np.abs(r).max())) < LAYER_TOL[_lib.MATH_TC_BF16X3]


def test_predict_after_uint8_input_uses_the_float_input():
    """predict() after a uint8 run must read ITS float32 input, not the uint8 buffer still bound from before."""
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=2)
    weights = random_weights(m.engine.expected_variables(), seed=5, detection_rich=True)
    m.set_weights_dict(weights)
    rng = np.random.default_rng(0)
    a_u8 = torch.from_numpy(rng.integers(0, 256, (2, 224, 320, 3), dtype=np.uint8)).cuda()
    b = rng.random((2, 224, 320, 3), dtype=np.float32)
    fresh = [t.copy() for t in m.predict(b)]
    m.predict_device_u8(a_u8)
    again = m.predict(b)
    for p, q in zip(fresh, again):
        np.testing.assert_array_equal(p, q)
    # ... and an externally bound input buffer (double-buffered ingest) gives the same heads as the engine's own
    ext = torch.from_numpy(b).cuda()
    m.engine.bind_input(ext)
    via_ext = [t.clone() for t in m.engine.run(2)]
    for p, q in zip(fresh, via_ext):
        np.testing.assert_array_equal(p, q.cpu().numpy())
    m.engine.unbind_input()


def test_errors():
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=2)
    with pytest.raises(_lib.K2YError):
        m.predict(np.zeros((1, 224, 320, 3), np.float32))            # weights not loaded
    with pytest.raises(ValueError):
        m.set_weights_dict({})
    weights = random_weights(m.engine.expected_variables(), seed=5)
    bad = dict(weights)
    bad["conv1"] = {"kernel": np.zeros((3, 3, 3, 8), np.float32)}
    with pytest.raises(ValueError):
        m.set_weights_dict(bad)
    m.set_weights_dict(weights)
    with pytest.raises(ValueError):
        m.predict(np.zeros((1, 100, 100, 3), np.float32))
    with pytest.raises(ValueError):
        m.predict_device(torch.zeros((3, 224, 320, 3), device="cuda"))  # > max_batch
