"""Quick GPU probe: FP64 peak, E_cost timings at the BASELINE sizes."""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96  # noqa: E402
from meanfieldvargp_b200 import _lib  # noqa: E402

lib = _lib.require_device()
tf = ctypes.c_double()
_lib.check(lib.mfvgp_fp64_peak(0, ctypes.byref(tf)), "peak")
print(f"fp64 DFMA peak: {tf.value:.2f} TFLOP/s")

sizes = [(500, 16), (2048, 257), (8192, 513)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for D, M in sizes:
    g = synthetic_l96(D, M, seed=8192)
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    for _ in range(3):
        F, grad = fe.E_cost(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        F, grad = fe.E_cost(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cells = D * (M - 1)
    print(f"L96 D={D} M={M}: {ms*1e3:.1f} us/eval, {4620*cells/ms/1e9:.2f} TFLOP/s algorithmic, "
          f"{96*cells/ms/1e6:.1f} GB/s algorithmic, F={F.item():.10g}")
