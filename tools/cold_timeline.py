"""CUPTI timeline of E_cost (L96 D=500, device pointers) with the L2 flushed before every call --
the protocol of bench.py's `value` -- and back to back."""
import sys
from pathlib import Path
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for _ in range(5):
    fe.E_cost(x)
torch.cuda.synchronize()
for cold in (True, False):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(8):
            if cold:
                flush.zero_()
            fe.E_cost(x)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.end)
    rows = [(e.name.split("(")[0].split("::")[-1][:34], e.time_range.start, e.time_range.end) for e in ev]
    mains = [i for i, r in enumerate(rows) if "coop" in r[0]]
    i0 = mains[5]
    i1 = mains[6]
    t0 = rows[i0][1]
    print("COLD (L2 flushed)" if cold else "WARM (back to back)")
    for name, st, en in rows[i0:i1 + 1]:
        print(f"  {name:36s} start {st - t0:8.2f}  end {en - t0:8.2f}  dur {en - st:7.2f}")
