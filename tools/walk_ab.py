"""Walk kernel A/B at D=8192: kernel time (library events) and evaluation time for L = 512 and
L = 4096 under the environment of the caller (MFVGP_LIB, MFVGP_FTILE, MFVGP_WALK_SCHED ...), with
F and the gradient sum for comparing variants.  `python tools/walk_ab.py [label]`."""
import ctypes, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200 import _lib
from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96_device
lib = _lib.load()
label = sys.argv[1] if len(sys.argv) > 1 else "-"
for M in (513, 4097):
    g = synthetic_l96_device(torch, 8192, M, seed=8192)
    fe = make_free_energy(g)
    for _ in range(3):
        F, grad = fe.E_cost(g["x0"])
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    lib.mfvgp_timing(fe._h, 1, ctypes.byref(tot), ctypes.byref(cnt))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F, grad = fe.E_cost(g["x0"])
    e1.record(); torch.cuda.synchronize()
    lib.mfvgp_timing(fe._h, 0, ctypes.byref(tot), ctypes.byref(cnt))
    d = fe.describe()
    print(f"{label:12s} L={M-1:5d} te={d['te']} nruns={d['nruns']:>4s} kernel {1e3*tot.value/cnt.value:8.1f} us  "
          f"eval {1e3*e0.elapsed_time(e1)/10:8.1f} us  F={F.item():.15e} gsum={grad.sum().item():.15e} "
          f"status {fe.check_last()}", flush=True)
    del fe, grad, g
    torch.cuda.empty_cache()
