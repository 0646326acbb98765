"""SCG on the long-horizon problem (L96 D=8192, M=4097, one GPU): time per iteration as the run
goes on, rejected steps, and how many intervals the guard flags."""
import contextlib, io, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200.synthetic import synthetic_l96_device, make_free_energy
g = synthetic_l96_device(torch, 8192, 4097, seed=8192)
fe = make_free_energy(g)
x = g["x0"]
with contextlib.redirect_stdout(io.StringIO()):
    fe.find_minimum(x, max_iter=10)
for it in (12, 20, 30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        xo, fx, st = fe.find_minimum(x, max_iter=it)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    F, _ = fe.E_cost(xo)
    stt = fe.check_last()
    info = fe.refine_info()
    rej = (3 * st["nit"] + 1 - st["func_eval"]) // 2
    print(f"max_iter={it}: nit {st['nit']} evals {st['func_eval']} rejected {rej} {dt*1e3:.1f} ms "
          f"{st['nit']/dt:.1f} it/s; at the end: status {stt}, flagged intervals "
          f"{int((info > 0).any(axis=1).sum())} of {info.shape[0]}", flush=True)
