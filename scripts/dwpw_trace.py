"""Device-side timeline of the fused depthwise->pointwise kernel, one block at a time (K2Y_TC_TRACE=1)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads as wl
from k210_yolo_framework_b200 import yolonet
cfg = wl.CONFIGS[2]
m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=32)
m.set_weights_dict(wl.bench_weights(cfg, m.engine.expected_variables()))
m.engine.set_use_graph(False)
x = torch.rand((32, 224, 320, 3), device="cuda")
os.environ["K2Y_DWPW"] = "1"
for blk in [int(a) for a in sys.argv[1:]] or [1, 3, 5, 7]:
    os.environ["K2Y_DWPW_MASK"] = hex(1 << (blk - 1))
    os.environ.pop("K2Y_TC_TRACE", None)
    m.predict_device(x)
    torch.cuda.synchronize()
    os.environ["K2Y_TC_TRACE"] = "1"
    print(f"==== block {blk}", file=sys.stderr, flush=True)
    m.predict_device(x)
    torch.cuda.synchronize()
