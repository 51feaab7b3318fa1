"""Writes tests/golden/vars_cfgN.json: {keras_layer: {var: shape}} of every BASELINE workload's network, taken from the
native graph builder (k2y_net_create needs no GPU).  bench.py's reference arm builds its seeded weights from these
files so that it never has to load this repo's native library; tests/test_graph_builder.py keeps them equal to the builder."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads as wl  # noqa: E402
from k210_yolo_framework_b200 import yolonet  # noqa: E402


def main():
    for cfg in wl.CONFIGS.values():
        h, w = cfg["in_hw"]
        model, _ = getattr(yolonet, cfg["model"])([h, w, 3], 3, cfg["classes"], alpha=cfg["alpha"], max_batch=1)
        exp = model.engine.expected_variables()
        path = os.path.join(wl.GOLDEN, f"vars_{cfg['name']}.json")
        with open(path, "w") as fh:
            json.dump({k: {v: list(s) for v, s in vs.items()} for k, vs in exp.items()}, fh, separators=(",", ":"))
        print(path, len(exp), "layers")


if __name__ == "__main__":
    main()
