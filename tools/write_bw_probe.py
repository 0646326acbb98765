"""Pure-write HBM bandwidth on this GPU (torch fill_ / zero_ of 6 GB, aligned and with the odd row
pitch of the posterior outputs) beside mfvgp_reconstruct at D=8192, M=4097, num=20."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
D, P = 8192, 81961
a = torch.empty((D, P), dtype=torch.float64, device="cuda")
ms = t(lambda: a.fill_(1.5)); print(f"fill_ [D][P] contiguous: {a.numel()*8/ms*1e-6:.0f} GB/s")
ms = t(lambda: a.zero_()); print(f"zero_ (memset): {a.numel()*8/ms*1e-6:.0f} GB/s")
b = torch.empty((D, P + 3), dtype=torch.float64, device="cuda")
ms = t(lambda: b[:, :P].fill_(1.5)); print(f"fill_ rows of P in pitch P+3 (32B-aligned rows): {D*P*8/ms*1e-6:.0f} GB/s")
c = torch.empty_like(a)
ms = t(lambda: c.copy_(a)); print(f"copy_ (read+write bytes): {2*a.numel()*8/ms*1e-6:.0f} GB/s")
