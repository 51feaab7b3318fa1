import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_weights():
    from k210_yolo_framework_b200.yolonet import load_npz_weights
    return load_npz_weights(os.path.join(GOLDEN, "yolo_mobilev1_075_voc_weights.npz"))


@pytest.fixture(scope="session")
def voc_anchors():
    return np.load(os.path.join(GOLDEN, "voc_anchor.npy"))


@pytest.fixture(scope="session")
def dog_u8():
    return np.load(os.path.join(GOLDEN, "dog_u8.npy"))


@pytest.fixture(scope="session")
def people_u8():
    return np.load(os.path.join(GOLDEN, "people_u8.npy"))


@pytest.fixture(scope="session")
def dog_heads():
    z = np.load(os.path.join(GOLDEN, "dog_heads.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def dog_golden():
    import json
    with open(os.path.join(GOLDEN, "dog_golden.json")) as fh:
        return json.load(fh)
