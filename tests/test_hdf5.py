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
