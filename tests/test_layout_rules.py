"""Scope rules of the build: the product never touches the oracle, and has no CPU fallback."""
import os
import re

from conftest import ROOT


def _py_files(d):
    for base, _dirs, files in os.walk(d):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                yield os.path.join(base, f)


def test_product_never_imports_oracle():
    bad = []
    for path in list(_py_files(os.path.join(ROOT, "k210_yolo_framework_b200"))) + [os.path.join(ROOT, "keras_inference.py")]:
        src = open(path).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/_ref" in src:
            bad.append(path)
    assert not bad, f"product files reference the oracle: {bad}"


def test_oracle_is_labelled_test_infrastructure():
    for f in ("__init__.py", "keras_ref.py", "decode_ref.py", "region_c.py"):
        assert "TEST INFRASTRUCTURE" in open(os.path.join(ROOT, "oracle", f)).read()


def test_missing_gpu_fails_loudly():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from k210_yolo_framework_b200 import K2YError, yolo_mobilev1, KerasDetector
    m, _ = yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75)
    with pytest.raises(K2YError):
        m.engine.set_weights({})
    with pytest.raises(K2YError):
        m.predict(np.zeros((1, 224, 320, 3), np.float32))
    with pytest.raises(K2YError):
        KerasDetector(np.zeros((2, 3, 2)), [224, 320], [7, 10, 14, 20], 20)
