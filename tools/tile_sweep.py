import ctypes, os, sys, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from helpers import make_free_energy, synthetic_l96_device
    from meanfieldvargp_b200 import _lib
    g = synthetic_l96_device(torch, 8192, 1025, seed=8192)
    fe = make_free_energy(g); lib = _lib.load()
    for _ in range(3): fe.E_cost(g["x0"])
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    lib.mfvgp_timing(fe._h, 1, ctypes.byref(tot), ctypes.byref(cnt))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fe.E_cost(g["x0"])
    e1.record(); torch.cuda.synchronize()
    lib.mfvgp_timing(fe._h, 0, ctypes.byref(tot), ctypes.byref(cnt))
    print(f"{os.environ.get('MFVGP_FTILE','default'):>8} fuse={os.environ.get('MFVGP_FUSE','0')}: main kernel {tot.value/cnt.value*1e3:8.1f} us, eval {e0.elapsed_time(e1)/10*1e3:8.1f} us")
else:
    for cfg in ["32,4", "32,8", "64,1", "64,2", "64,4", "128,1", "128,2"]:
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MFVGP_FTILE=cfg, MFVGP_FUSE="1"))
    subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ))
