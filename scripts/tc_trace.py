"""Device-side timeline of the tensor-core conv kernel (K2Y_TC_TRACE=1) on a few layer shapes of cfg 2."""
import ctypes, os, sys, time
os.environ["K2Y_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from k210_yolo_framework_b200 import _lib
from k210_yolo_framework_b200._lib import check, lib

def run(B, H, W, C0, C1, up0, Cout, k, math=None, reps=2):
    math = int(os.environ.get("K2Y_TRACE_MATH", "1")) if math is None else math
    rng = np.random.default_rng(0)
    x0 = torch.randn((B, H, W, C0), device="cuda")
    hh, ww = (H * 2, W * 2) if up0 else (H, W)
    x1 = torch.randn((B, hh, ww, C1), device="cuda") if C1 else None
    out = torch.empty((B, hh, ww, Cout), device="cuda")
    kern = (rng.normal(0, 1, (k, k, C0 + C1, Cout)) / np.sqrt(k * k * (C0 + C1))).astype(np.float32)
    sc, sh = np.ones(Cout, np.float32), np.zeros(Cout, np.float32)
    st = torch.cuda.current_stream()
    for r in range(reps):
        print(f"--- {B}x{hh}x{ww} C={C0}+{C1} -> {Cout} k={k} rep {r}", file=sys.stderr, flush=True)
        check(lib.k2y_conv2d(x0.data_ptr(), x1.data_ptr() if x1 is not None else None, None, out.data_ptr(), kern.ctypes.data,
                             sc.ctypes.data, sh.ctypes.data, B, hh, ww, C0, C1, int(up0), Cout, k, 1, 0, 1, 0.1, math,
                             ctypes.c_void_p(st.cuda_stream)))

MATH = int(os.environ.get("K2Y_TRACE_MATH", "1"))
which = sys.argv[1:] or ["h1o", "pw1", "pw3", "pw7", "h13"]
if "h1o" in which: run(32, 7, 10, 192, 0, 0, 75, 1)      # head1_out
if "pw1" in which: run(32, 112, 160, 24, 0, 0, 48, 1)    # pw1
if "pw3" in which: run(32, 56, 80, 96, 0, 0, 96, 1)      # pw3
if "pw7" in which: run(32, 14, 20, 384, 0, 0, 384, 1)    # pw7
if "h13" in which: run(32, 7, 10, 768, 0, 0, 192, 3)     # head1 3x3
if "h23" in which: run(32, 7, 10, 128, 384, 1, 128, 3)   # head2 3x3: up(lateral 128) || x1 (384) -> 128
if "pw1p" in which: run(1, 1, 286720, 48, 0, 0, 96, 1)   # conv_pw_1 in pixel-pair form: [M/2][48] x blockdiag -> [M/2][96]
if "pw2" in which: run(32, 56, 80, 48, 0, 0, 96, 1)      # pw2
