"""k210_yolo_framework_b200 — B200-native YOLOv3 inference hot path of zhen8838/K210_Yolo_framework.

Importing the package loads libk210yolo_b200.so (hand-written sm_100a CUDA behind a C-ABI).  There is no
CPU fallback: a missing library is an ImportError, a missing GPU a ``K2YError`` at the first compute call.
"""
from ._lib import K2YError, MATH_FP32_SIMT, MATH_TC_3XTF32, MATH_TC_TF32, MATH_TC_BF16X3, cuda_available  # noqa: F401
from .yolonet import yolo_mobilev1, yolo_mobilev2, tiny_yolo, yolo, YoloEngine, YoloModel  # noqa: F401
from .helper import Helper  # noqa: F401
from .detect import KerasDetector, RegionDetector  # noqa: F401

__all__ = ["yolo_mobilev1", "yolo_mobilev2", "tiny_yolo", "yolo", "YoloEngine", "YoloModel", "Helper",
           "KerasDetector", "RegionDetector", "K2YError", "cuda_available", "MATH_FP32_SIMT", "MATH_TC_3XTF32",
           "MATH_TC_TF32", "MATH_TC_BF16X3"]
