"""Kernel timeline (CUPTI through torch.profiler) of a device SCG run or of back-to-back E_cost
calls on L96 D=500: start offset, duration and gap of every kernel of a few iterations."""
import contextlib, io, sys
from pathlib import Path
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96  # noqa: E402
what = sys.argv[1] if len(sys.argv) > 1 else "scg"
D, M = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (500, 16)
g = synthetic_l96(D, M, seed=8192)
fe = make_free_energy(g)
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
with contextlib.redirect_stdout(io.StringIO()):
    fe.find_minimum(x, max_iter=10)
    for _ in range(5):
        fe.E_cost(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with contextlib.redirect_stdout(io.StringIO()):
        if what == "scg":
            fe.find_minimum(x, max_iter=30)
        else:
            for _ in range(30):
                fe.E_cost(x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
rows = [(e.name.split("(")[0].split("::")[-1][:34], e.time_range.start - t0, e.time_range.end - e.time_range.start) for e in ev]
n = len(rows)
lo = max(0, n // 2 - 12)
prev_end = None
for name, st, du in rows[lo:lo + 26]:
    gap = "" if prev_end is None else f"gap {st - prev_end:6.2f}"
    print(f"{name:36s} start {st:9.2f} us  dur {du:6.2f} us  {gap}")
    prev_end = st + du
tot = rows[-1][1] + rows[-1][2] - rows[0][1]
busy = sum(r[2] for r in rows)
print(f"kernels {n}, span {tot:.1f} us, busy {busy:.1f} us ({100 * busy / tot:.0f} %)")
