"""Seeded Keras-default initialisation for the four networks (synthetic benchmark weights).

BASELINE.json's benchmark configs use random-init weights (no checkpoints offline).  Keras defaults,
restated: Conv2D / DepthwiseConv2D kernels ``glorot_uniform`` (limit sqrt(6/(fan_in+fan_out)), with
fan_in = kh*kw*cin, fan_out = kh*kw*cout; depthwise: fan_in = 9*C... computed on the (3,3,C,1) shape as
Keras does: receptive 9, fan_in = 9*C, fan_out = 9*1), bias zeros, BatchNormalization gamma=1, beta=0,
moving_mean=0, moving_variance=1.

``detection_rich=True`` additionally perturbs BN statistics and the final 1x1 biases with seeded noise so
that a controlled fraction of (box, class) scores passes the 0.7 threshold — with pure default init
sigmoid(cls)*sigmoid(conf) ~ 0.25 and NMS would see no candidates.

The result is the ``{keras_layer: {var: ndarray}}`` dict both the CUDA engine and the oracle consume.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

Weights = Dict[str, Dict[str, np.ndarray]]


def random_weights(expected: Dict[str, Dict[str, tuple]], seed: int = 0, detection_rich: bool = False,
                   head_bias: float = 1.5, head_bias_std: float = 2.0) -> Weights:
    rng = np.random.default_rng(seed)
    out: Weights = {}
    for layer, vars_ in expected.items():
        out[layer] = {}
        for var, shape in vars_.items():
            if var == "kernel":
                kh, kw, cin, cout = shape
                limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
                arr = rng.uniform(-limit, limit, shape)
            elif var == "depthwise_kernel":
                kh, kw, c, m = shape
                limit = np.sqrt(6.0 / (kh * kw * c + kh * kw * m))
                arr = rng.uniform(-limit, limit, shape)
            elif var in ("gamma", "moving_variance"):
                arr = np.ones(shape)
                if detection_rich:
                    arr = arr * rng.uniform(0.8, 1.25, shape)
            else:  # bias, beta, moving_mean
                arr = np.zeros(shape)
                if detection_rich and var in ("beta", "moving_mean"):
                    arr = rng.normal(0.0, 0.05, shape)
            out[layer][var] = arr.astype(np.float32)
    if detection_rich:
        # bias the conf/class logits of the final 1x1 convs (the only layers with a bias)
        for layer, vars_ in out.items():
            if "bias" in vars_:
                n = vars_["bias"].shape[0]
                vars_["bias"] = (head_bias + rng.normal(0.0, head_bias_std, n)).astype(np.float32)
                # make the final 1x1 kernel strong enough that logits vary across the grid
                vars_["kernel"] = (vars_["kernel"] * 8.0).astype(np.float32)
    return out
