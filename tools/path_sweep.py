"""Times E_cost of L96 D=8192 for the E_cost paths / tiles (CUDA events).
usage: python tools/path_sweep.py [M] [cfg ...]   cfg = path[:TE,TN]"""
import ctypes, os, sys, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from helpers import make_free_energy, synthetic_l96_device
    from meanfieldvargp_b200 import _lib
    M = int(sys.argv[2]); D = int(os.environ.get("SWEEP_D", "8192"))
    g = synthetic_l96_device(torch, D, M, seed=8192)
    fe = make_free_energy(g); lib = _lib.load()
    for _ in range(3): F, _g = fe.E_cost(g["x0"])
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    lib.mfvgp_timing(fe._h, 1, ctypes.byref(tot), ctypes.byref(cnt))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fe.E_cost(g["x0"])
    e1.record(); torch.cuda.synchronize()
    lib.mfvgp_timing(fe._h, 0, ctypes.byref(tot), ctypes.byref(cnt))
    print(f"{Path(os.environ.get('MFVGP_LIB','base')).stem:>16} {os.environ.get('MFVGP_PATH','default'):>6} {os.environ.get('MFVGP_FTILE','auto'):>7} D={D} M={M}: "
          f"main kernel {tot.value/cnt.value*1e3:9.1f} us, eval {e0.elapsed_time(e1)/10*1e3:9.1f} us, F={F.item():.15g}", flush=True)
else:
    M = sys.argv[1] if len(sys.argv) > 1 else "1025"
    cfgs = sys.argv[2:] or ["split", "walk", "walk:128,8", "walk:128,16", "walk:128,32", "walk:128,64", "walk:64,16"]
    for cfg in cfgs:
        path, _, tile = cfg.partition(":")
        env = dict(os.environ, MFVGP_PATH=path)
        if tile: env["MFVGP_FTILE"] = tile
        subprocess.run([sys.executable, __file__, "child", M], env=env)
