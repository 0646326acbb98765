"""Walk-kernel run schedules at D=8192: kernel time (library events) for L = 512 (the shard of an
8-GPU run) and L = 4096, MFVGP_WALK_SCHED variants."""
import ctypes, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200 import _lib
from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96_device
lib = _lib.load()
scheds = sys.argv[1:] or ["uniform", "2,2,32", "2,2,64", "1.5,2,32", "3,2,32", "2,3,24", "2,1,32", "2.5,2,48"]
for M in (513, 4097):
    g = synthetic_l96_device(torch, 8192, M, seed=8192)
    ref = None
    for sc in scheds:
        os.environ["MFVGP_WALK_SCHED"] = sc
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
        cs = float(grad.sum().item())
        ref = ref if ref is not None else (float(F.item()), cs)
        d = fe.describe()
        print(f"L={M-1} sched={sc:10s} nruns={d['nruns']:>4s} tn0={d['tn']:>3s} kernel {1e3*tot.value/cnt.value:8.1f} us  eval {1e3*e0.elapsed_time(e1)/10:8.1f} us  "
              f"dF={abs(F.item()-ref[0])/abs(ref[0]):.1e} dsum={abs(cs-ref[1])/abs(ref[1]):.1e}", flush=True)
        del fe, grad
    del g
    torch.cuda.empty_cache()
