"""Where the time of FreeEnergy.E_cost(numpy) goes (L96 D=500, M=16), microseconds."""
import ctypes, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96
from meanfieldvargp_b200 import _lib
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
lib = _lib.load()
x_pin = torch.tensor(g["x0"], dtype=torch.float64).pin_memory(); x = x_pin.numpy()
n = x.size
def t(f, reps=2000):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e6
print("E_cost(numpy pinned x)      %.1f us" % t(lambda: fe.E_cost(x)))
print("E_cost(numpy pageable x)    %.1f us" % t(lambda: fe.E_cost(g["x0"])))
print("  _sync_params              %.1f us" % t(fe._sync_params))
print("  pinned_empty(n+8)         %.1f us" % t(lambda: _lib.pinned_empty(n + 8)))
buf = _lib.pinned_empty(n + 8)
print("  C call (H2D+kernels+D2H+sync) %.1f us" % t(lambda: lib.mfvgp_ecost_host_packed(fe._h, x.ctypes.data, 1, buf.ctypes.data)))
print("  x.ctypes.data             %.2f us" % t(lambda: x.ctypes.data))
print("  status parse              %.2f us" % t(lambda: int(buf[n + 4:n + 5].view(np.int32)[0])))
print("  terms dict                %.2f us" % t(lambda: {"F": buf[n], "E0": buf[n+1], "Esde": buf[n+2], "Eobs": buf[n+3]}))
print("  buf[:n]                   %.2f us" % t(lambda: buf[:n]))
xd = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
out = torch.empty(8, dtype=torch.float64, device="cuda"); gr = torch.empty(n, dtype=torch.float64, device="cuda")
def dev_sync():
    lib.mfvgp_ecost(fe._h, xd.data_ptr(), 1, out.data_ptr(), gr.data_ptr(), out.data_ptr() + 32, 0)
    torch.cuda.synchronize()
print("  device-only eval + sync   %.1f us" % t(dev_sync))
