"""The native graph builders (k2y_net_create) against the oracle's restatement of the reference builders:
same Keras layer names, same variable shapes, same output grids — for all four model_defs."""
import numpy as np
import pytest

from k210_yolo_framework_b200 import yolonet
from k210_yolo_framework_b200.weights import random_weights
from oracle import keras_ref

CASES = [("yolo_mobilev1", 0.75, 20, (64, 96)), ("yolo_mobilev1", 1.0, 20, (64, 64)), ("yolo_mobilev1", 0.5, 4, (64, 64)),
         ("yolo_mobilev2", 1.0, 20, (64, 96)), ("yolo_mobilev2", 0.5, 20, (64, 64)), ("yolo_mobilev2", 0.75, 7, (64, 64)),
         ("tiny_yolo", 1.0, 20, (96, 96)), ("yolo", 1.0, 80, (64, 64))]


@pytest.mark.parametrize("model_def,alpha,classes,hw", CASES)
def test_graph_matches_oracle_structure(model_def, alpha, classes, hw):
    m, w = getattr(yolonet, model_def)([hw[0], hw[1], 3], 3, classes, alpha=alpha, max_batch=2)
    exp = m.engine.expected_variables()
    weights = random_weights(exp, seed=1)
    x = np.random.default_rng(0).random((1, hw[0], hw[1], 3), dtype=np.float32)
    # the oracle indexes weights by the names/shapes the reference builders would create: any mismatch raises
    heads = keras_ref.forward(model_def, weights, x, alpha=alpha)
    assert [h.shape[1:] for h in heads] == list(m.engine.out_shapes)
    assert m.output_shapes == [(None, h, ww, c) for h, ww, c in m.engine.out_shapes]
    assert w.output_shapes == [(None, h, ww, 3, 5 + classes) for h, ww, _ in m.engine.out_shapes]
    n_out = 2 if model_def != "yolo" else 3
    assert len(heads) == n_out
    for l, (h, ww, c) in enumerate(m.engine.out_shapes):
        assert (h, ww, c) == (hw[0] // 32 * 2 ** l, hw[1] // 32 * 2 ** l, 3 * (5 + classes))


def test_parameter_counts():
    def count(m):
        return sum(int(np.prod(s)) for v in m.engine.expected_variables().values() for s in v.values())
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75)
    assert count(m) == 3874150            # == asset/yolo_model.h5
    assert m.engine.out_shapes == [(7, 10, 75), (14, 20, 75)]
    m, _ = yolonet.yolo([416, 416, 3], 3, 80)
    assert m.engine.out_shapes == [(13, 13, 255), (26, 26, 255), (52, 52, 255)]
    assert abs(count(m) - 62.0e6) < 0.1e6  # canonical YOLOv3: 61.9 M + BN statistics


def test_keras_auto_names_of_heads():
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75)
    names = [L.name.decode() for L in m.engine.layers()]
    assert names[0] == "conv1" and names[1] == "conv_dw_1" and names[2] == "conv_pw_1"
    assert names[-5:] == ["conv2d", "conv2d_1", "conv2d_2", "conv2d_3", "conv2d_4"]  # as stored in asset/yolo_model.h5
    bns = [L.bn_name.decode() for L in m.engine.layers()][-5:]
    assert bns == ["batch_normalization", "", "batch_normalization_1", "batch_normalization_2", ""]


def test_auto_name_remap_for_offset_files():
    m, _ = yolonet.yolo_mobilev1([64, 64, 3], 3, 20, alpha=0.75)
    exp = m.engine.expected_variables()
    w = random_weights(exp, seed=3)
    shifted = {}
    for k, v in w.items():
        if k.startswith("conv2d"):
            i = int(k.split("_")[1]) if "_" in k else 0
            shifted[f"conv2d_{i + 5}"] = v
        elif k.startswith("batch_normalization"):
            i = int(k.rsplit("_", 1)[1]) if k != "batch_normalization" else 0
            shifted[f"batch_normalization_{i + 3}"] = v
        else:
            shifted[k] = v
    back = yolonet._rename_auto_named(shifted, exp)
    for k in exp:
        for var in exp[k]:
            assert back[k][var] is w[k][var]
