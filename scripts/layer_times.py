"""Per-launch times of one cfg2 step (batch 32) for the tensor-core arithmetic modes, side by side.
CUDA events between launches (k2y_net_profile), L2 flushed before every pass, mean of 5 passes."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from k210_yolo_framework_b200 import _lib, yolonet  # noqa: E402
from k210_yolo_framework_b200.weights import random_weights  # noqa: E402

BATCH = int(os.environ.get("BATCH", "32"))
m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=BATCH)
m.set_weights_dict(random_weights(m.engine.expected_variables(), seed=0, detection_rich=True))
x = torch.rand((BATCH, 224, 320, 3), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cols = {}
for mode in (_lib.MATH_TC_3XTF32, _lib.MATH_TC_BF16X3):
    m.engine.set_math(mode)
    for _ in range(3):
        m.predict_device(x)
    acc = None
    for _ in range(5):
        flush.zero_()
        prof = m.engine.profile(BATCH)
        acc = prof if acc is None else [dict(a, ms=a["ms"] + b["ms"]) for a, b in zip(acc, prof)]
    cols[mode] = [(a["name"], a["ms"] / 5 * 1e3) for a in acc]
print(f"{'layer':<28}{'3xtf32 us':>12}{'bf16x3 us':>12}")
for (n, a), (_, b) in zip(cols[_lib.MATH_TC_3XTF32], cols[_lib.MATH_TC_BF16X3]):
    print(f"{n:<28}{a:12.1f}{b:12.1f}")
print(f"{'sum':<28}{sum(a for _, a in cols[_lib.MATH_TC_3XTF32]):12.1f}{sum(b for _, b in cols[_lib.MATH_TC_BF16X3]):12.1f}")
