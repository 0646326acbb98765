"""Device-resident E_cost timing on the north-star size (L96 D=500, M=16)."""
import ctypes, os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96
from meanfieldvargp_b200 import _lib
D = int(os.environ.get("SB_D", "500")); M = int(os.environ.get("SB_M", "16"))
g = synthetic_l96(D, M, seed=8192)
fe = make_free_energy(g)
lib = _lib.load()
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
for _ in range(10): F, grad = fe.E_cost(x)
torch.cuda.synchronize()
# (1) pure GPU time of one evaluation: the GPU is kept busy while the CPU enqueues
gpu = []
for _ in range(30):
    torch.cuda._sleep(2_000_000)          # ~1 ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fe.E_cost(x); e1.record()
    torch.cuda.synchronize()
    gpu.append(e0.elapsed_time(e1) * 1e3)
gpu.sort()
# (2) CPU cost of the Python call (asynchronous enqueue)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): fe.E_cost(x)
cpu_py = (time.perf_counter() - t0) / 2000 * 1e6
torch.cuda.synchronize()
# (3) CPU cost of the bare C-ABI call on preallocated device buffers
n = x.numel()
out = torch.empty(4, dtype=torch.float64, device="cuda"); st = torch.empty(1, dtype=torch.int32, device="cuda")
gr = torch.empty(n, dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
args = (fe._h, x.data_ptr(), 1, out.data_ptr(), gr.data_ptr(), st.data_ptr(), stream)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): lib.mfvgp_ecost(*args)
cpu_c = (time.perf_counter() - t0) / 2000 * 1e6
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): lib.mfvgp_ecost(*args)
torch.cuda.synchronize()
thr_c = (time.perf_counter() - t0) / 2000 * 1e6
l0 = fe.kernel_launches; fe.E_cost(x); l1 = fe.kernel_launches
x_h = g["x0"]
for _ in range(10): fe.E_cost(x_h)
t0 = time.perf_counter()
for _ in range(300): fe.E_cost(x_h)
e2e = (time.perf_counter() - t0) / 300 * 1e6
print(f"D={D} M={M} COOP={os.environ.get('MFVGP_COOP','1')} PATH={os.environ.get('MFVGP_PATH','-')}: GPU/eval median {gpu[15]:.1f} us (min {gpu[0]:.1f}), "
      f"CPU/call python {cpu_py:.1f} us, C-ABI enqueue {cpu_c:.1f} us, C-ABI back-to-back {thr_c:.1f} us/eval, launches/eval {l1-l0}, e2e(numpy) {e2e:.1f} us")
