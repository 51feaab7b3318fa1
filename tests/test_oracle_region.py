"""numpy restatement of region_layer.c vs the compiled, unmodified reference (oracle/_ref)."""
import numpy as np
import pytest

from oracle import region_c

ANCH = [0.76120044, 0.57155991, 0.6923348, 0.88535553, 0.47163042, 0.34163313]

pytestmark = pytest.mark.skipif(not region_c.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("seed,w,h,classes,thr,nms,net,img", [
    (0, 10, 7, 20, 0.3, 0.3, (320, 224), (320, 224)),
    (1, 20, 14, 20, 0.2, 0.45, (320, 224), (320, 224)),
    (2, 6, 5, 3, 0.1, 0.3, (320, 224), (320, 224)),
    (3, 10, 7, 20, 0.3, 0.3, (320, 224), (499, 374)),   # letterbox correction active
    (4, 8, 8, 5, 0.25, 0.5, (256, 256), (200, 300)),
])
def test_numpy_restatement_matches_compiled_reference(seed, w, h, classes, thr, nms, net, img):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 2.0, (3, 5 + classes, h, w)).astype(np.float32)
    x[:, 4] += 2.0
    r = region_c.RegionLayerRef(w, h, 3 * (5 + classes), net[0], net[1], ANCH, thr, nms, image_width=img[0], image_height=img[1])
    ref = r.run(x)
    got, probs, boxes = region_c.region_layer_np(x, w, h, ANCH, thr, nms, net[0], net[1], img[0], img[1])
    assert len(ref) > 0
    assert [t[:5] for t in got] == [t[:5] for t in ref]
    np.testing.assert_allclose([t[5] for t in got], [t[5] for t in ref], atol=1e-6)
    assert ((r.probs() > 0) == (probs > 0)).all()
    np.testing.assert_allclose(probs, r.probs(), atol=1e-6)
    np.testing.assert_allclose(boxes, r.boxes(), rtol=1e-5, atol=1e-6)


def test_empty_result_when_threshold_high():
    x = np.zeros((3, 25, 7, 10), np.float32)
    r = region_c.RegionLayerRef(10, 7, 75, 320, 224, ANCH, 0.9, 0.3)
    assert r.run(x) == []
    assert region_c.region_layer_np(x, 10, 7, ANCH, 0.9, 0.3, 320, 224)[0] == []


def test_expf_restatement_is_libm_bit_for_bit():
    """oracle.region_c.expf_glibc (the algorithm csrc/detect.cu restates for the REGION_C dialect) against this host's
    libm expf — what the compiled reference region_layer.c calls — over normals, the subnormal tail and the limits."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(0, 4, 20000), rng.uniform(-110, 95, 10000),
                        [0.0, -0.0, 88.7228, 88.73, -103.97, -104.0, -90.0, -100.0, 1e-30, -1e-30]]).astype(np.float32)
    want = np.array([libm.expf(float(v)) for v in x], np.float32)
    got = region_c.expf_glibc(x)
    assert (want.view(np.uint32) == got.view(np.uint32)).all()
