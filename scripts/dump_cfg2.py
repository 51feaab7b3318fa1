import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from k210_yolo_framework_b200.pipeline import DetectionPipeline
from k210_yolo_framework_b200.weights import random_weights
ANCH2 = np.load(os.path.join(ROOT, "tests", "golden", "voc_anchor.npy"))
pipe = DetectionPipeline("yolo_mobilev1", (224, 320), ANCH2, 20, 0.75, 32, obj_thresh=0.7, iou_thresh=0.5)
w = random_weights(pipe.engine.expected_variables(), seed=0, detection_rich=True, head_bias=-0.2, head_bias_std=1.5)
pipe.engine.set_weights(w)
rng = np.random.default_rng(1)
x = rng.random((32, 224, 320, 3), dtype=np.float32); x[:, ::2, ::3] *= 0.5
pipe.engine.input_buffer.copy_(torch.from_numpy(x))
dets, counts = pipe.step_device(); torch.cuda.synchronize()
got = DetectionPipeline.records(dets.clone(), counts.clone())
np.savez(os.path.join(ROOT, "gpurun_out", "cfg2_dump.npz"), h0=pipe.engine.head_buffers[0][:2].cpu().numpy(), h1=pipe.engine.head_buffers[1][:2].cpu().numpy(),
         got0=np.array(got[0], dtype=np.float64), got1=np.array(got[1], dtype=np.float64))
print("dumped", len(got[0]))
