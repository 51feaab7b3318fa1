"""CPU oracle for the YOLOv3 inference hot path — TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference computes on the path
``keras_inference.py:main`` -> model builders -> decode -> per-class NMS
(/root/reference/keras_inference.py:75-135, models/yolonet.py, models/keras_mobilenet*.py,
tools/utils.py:524-547) and what the firmware's ``region_layer.c`` computes.

Rules (enforced by tests/test_layout_rules.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import anything under ``oracle/``;
  * the product package ``k210_yolo_framework_b200`` never imports it and has no CPU
    fallback — it raises when the CUDA library is missing.

Parity pinning status:
  * TensorFlow 1.14 (requirements.txt:3) cannot be installed here, and the reference has
    no tests, so the KERAS-dialect oracle is pinned only by the reference's shipped
    fixtures: ``asset/yolo_model.h5`` + ``data/dog.jpg`` must reproduce the two boxes drawn
    in ``asset/dog_res.jpg`` (class 11 "1.00", class 6 "0.8x").  tests/test_oracle_golden.py
    holds those known answers.  For the other three backbones (no shipped weights) the
    network restatement is "parity unpinned" beyond the layer semantics shared with
    yolo_mobilev1.
  * The REGION_C-dialect restatement (oracle/region_c.py) is pinned against the
    reference's own ``region_layer.c`` compiled unmodified into ``oracle/_ref/`` (see
    oracle/Makefile) on seeded inputs and on the dog.jpg head tensors.
"""
