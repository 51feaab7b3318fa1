"""Device-side timeline of the fused depthwise+pointwise launches (cfg2, batch 32): K2Y_TC_TRACE=1 on one eager pass."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from k210_yolo_framework_b200 import yolonet  # noqa: E402
from k210_yolo_framework_b200.weights import random_weights  # noqa: E402

m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=32)
m.set_weights_dict(random_weights(m.engine.expected_variables(), seed=0, detection_rich=True))
m.engine.set_use_graph(False)
x = torch.rand((32, 224, 320, 3), device="cuda")
for _ in range(3):
    m.predict_device(x)
torch.cuda.synchronize()
os.environ["K2Y_TC_TRACE"] = "1"
m.predict_device(x)
torch.cuda.synchronize()
