"""Minimal HDF5 reader vs the reference's shipped Keras model file (skipped where /root/reference is absent)."""
import json
import os

import numpy as np
import pytest

REF_H5 = "/root/reference/asset/yolo_model.h5"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference assets not present on this box")


def test_reads_every_dataset_like_the_committed_npz(golden_weights):
    from k210_yolo_framework_b200.hdf5_min import load_keras_weights
    w = load_keras_weights(REF_H5)
    assert sorted(w) == sorted(golden_weights)
    total = 0
    for layer in w:
        assert sorted(w[layer]) == sorted(golden_weights[layer])
        for var in w[layer]:
            assert w[layer][var].dtype == np.float32
            np.testing.assert_array_equal(w[layer][var], golden_weights[layer][var])
            total += w[layer][var].size
    assert total == 3874150


def test_attributes_and_groups():
    from k210_yolo_framework_b200.hdf5_min import File, Group
    f = File(REF_H5)
    assert f.root.keys() == ["model_weights"]
    assert f.root.attrs["keras_version"] == "2.2.4-tf"
    cfg = json.loads(f.root.attrs["model_config"])
    assert len(cfg["config"]["layers"]) == 100
    mw = f.root["model_weights"]
    assert isinstance(mw, Group) and len(mw.attrs["layer_names"]) == 100
    ds = f.root["model_weights/conv1/conv1/kernel:0"]
    assert ds.shape == (3, 3, 3, 24) and ds.read().shape == (3, 3, 3, 24)
    with pytest.raises(KeyError):
        f.root["model_weights/nope"]


def test_rejects_non_hdf5(tmp_path):
    from k210_yolo_framework_b200.hdf5_min import File, HDF5FormatError
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(HDF5FormatError):
        File(str(p))


def test_writer_round_trip(tmp_path, golden_weights):
    """hdf5_write.save_keras_weights -> hdf5_min.load_keras_weights: same layers, same order, same bits; both containers."""
    from k210_yolo_framework_b200.hdf5_min import File, load_keras_weights
    from k210_yolo_framework_b200.hdf5_write import save_keras_weights
    for full in (True, False):
        path = tmp_path / f"w{int(full)}.h5"
        save_keras_weights(str(path), golden_weights, full_model=full)
        back = load_keras_weights(str(path))
        assert sorted(back) == sorted(golden_weights)
        for layer, vars_ in golden_weights.items():
            assert set(back[layer]) == set(vars_)
            for var, arr in vars_.items():
                np.testing.assert_array_equal(back[layer][var], np.asarray(arr, np.float32))
        f = File(str(path))
        top = f.root["model_weights"] if full else f.root
        names = [n.decode() if isinstance(n, bytes) else str(n) for n in np.asarray(top.attrs["layer_names"]).ravel()]
        assert names == list(golden_weights)              # Keras' load_weights walks this attribute, in order
        wn = [n.decode() for n in np.asarray(top["conv1_bn"].attrs["weight_names"]).ravel()]
        assert wn == ["conv1_bn/gamma:0", "conv1_bn/beta:0", "conv1_bn/moving_mean:0", "conv1_bn/moving_variance:0"]


def test_writer_many_links_and_empty(tmp_path):
    """More links than one symbol-table leaf holds (64) and a zero-sized dataset."""
    from k210_yolo_framework_b200.hdf5_min import load_keras_weights
    from k210_yolo_framework_b200.hdf5_write import save_keras_weights
    rng = np.random.default_rng(0)
    w = {f"conv2d_{i}": {"kernel": rng.normal(size=(1, 1, 3, 2)).astype(np.float32)} for i in range(150)}
    w["empty"] = {"bias": np.zeros((0,), np.float32)}
    save_keras_weights(str(tmp_path / "many.h5"), w)
    back = load_keras_weights(str(tmp_path / "many.h5"))
    assert set(back) == set(w)
    for k in w:
        for v in w[k]:
            np.testing.assert_array_equal(back[k][v], w[k][v])


def test_fold_batchnorm_matches_conv_plus_bn():
    import torch
    import torch.nn.functional as F
    from k210_yolo_framework_b200.hdf5_write import fold_batchnorm
    rng = np.random.default_rng(1)
    w = {"c": {"kernel": rng.normal(size=(3, 3, 5, 7)).astype(np.float32)},
         "c_bn": {k: rng.uniform(0.5, 1.5, 7).astype(np.float32) for k in ("gamma", "beta", "moving_mean", "moving_variance")},
         "d": {"depthwise_kernel": rng.normal(size=(3, 3, 7, 1)).astype(np.float32)},
         "d_bn": {k: rng.uniform(0.5, 1.5, 7).astype(np.float32) for k in ("gamma", "beta", "moving_mean", "moving_variance")},
         "o": {"kernel": rng.normal(size=(1, 1, 7, 4)).astype(np.float32), "bias": rng.normal(size=4).astype(np.float32)}}
    f = fold_batchnorm(w, None, {"c": "c_bn", "d": "d_bn", "o": ""})
    x = torch.from_numpy(rng.normal(size=(2, 5, 9, 11)).astype(np.float32))

    def bn(t, p):
        g, b, m, v = (torch.from_numpy(p[k])[None, :, None, None] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        return (t - m) * torch.rsqrt(v + 1e-3) * g + b
    k = torch.from_numpy(w["c"]["kernel"]).permute(3, 2, 0, 1)
    y = bn(F.conv2d(x, k, padding=1), w["c_bn"])
    y2 = F.conv2d(x, torch.from_numpy(f["c"]["kernel"]).permute(3, 2, 0, 1), torch.from_numpy(f["c"]["bias"]), padding=1)
    assert float((y - y2).abs().max()) < 2e-5
    dk = torch.from_numpy(w["d"]["depthwise_kernel"]).permute(2, 3, 0, 1)
    z = bn(F.conv2d(y, dk, padding=1, groups=7), w["d_bn"])
    z2 = F.conv2d(y, torch.from_numpy(f["d"]["depthwise_kernel"]).permute(2, 3, 0, 1), torch.from_numpy(f["d"]["bias"]), padding=1, groups=7)
    assert float((z - z2).abs().max()) < 5e-5
    np.testing.assert_array_equal(f["o"]["bias"], w["o"]["bias"])
    np.testing.assert_array_equal(f["o"]["kernel"], w["o"]["kernel"])
