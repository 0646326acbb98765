import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, make_oracle, rel_err
import importlib.util
spec = importlib.util.spec_from_file_location("t", str(ROOT / "tests/test_gpu_random_sweep.py")); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = t._random_problem(seed)
o = make_oracle(g, adaptive=True); Fo, go = o.E_cost(g["x0"]); nev = np.array(o.neval)
print("oracle neval E:", nev[:, 0], "dS:", nev[:, 2])
for env in ({}, {"MFVGP_PATH": "walk"}, {"MFVGP_PATH": "split", "MFVGP_COOP": "0"}):
    os.environ.update(env)
    fe = make_free_energy(g)
    for k in env: os.environ.pop(k)
    F, grad = fe.E_cost(g["x0"]); i1 = fe.refine_info().copy()
    Fv = fe.E_cost(g["x0"], output_gradients=False); i2 = fe.refine_info().copy()
    print(env, "grad call E:", i1[:, 0], "dS:", i1[:, 1], "| value call E:", i2[:, 0], "dS:", i2[:, 1], "relF", abs(F-Fo)/abs(Fo), abs(Fv-Fo)/abs(Fo), "relg", rel_err(grad, go))
