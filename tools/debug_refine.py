import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, make_oracle, synthetic_l96, rel_err
D, M = 67, 12
g = synthetic_l96(D, M, seed=D + M)
x = g["x0"].copy()
rng = np.random.default_rng(M)
sp = x[D * (3 * M + 4):].reshape(D, 2 * M + 3)
cols = rng.choice(2 * M + 3, size=max(2, M // 6), replace=False)
sp[:, cols] = np.log(10.0 ** rng.uniform(-2.3, -1.0, size=(D, cols.size)))
fe = make_free_energy(g)
F, gr = fe.E_cost(x)
info = fe.refine_info()
o = make_oracle(g, adaptive=True)
Fo, go = o.E_cost(x)
nev = np.array(o.neval)
print("cols", cols)
print("gpu info E ", info[:,0]); print("scipy nev E", (nev[:,0]-21)//42)
print("gpu info dS", info[:,1]); print("scipy nev dS", (nev[:,2]-21)//42)
print("F", F, Fo, (F-Fo)/Fo, "terms", fe.last_terms)
mp = x[:D*(3*M+4)].reshape(D,-1); spv = np.exp(x[D*(3*M+4):].reshape(D,-1))
Es, dm, ds = fe.E_sde(mp, spv)
Eo, dmo, dso = o.E_sde(mp, spv)
print("Esde", Es, Eo, "per-interval dS err", [f"{rel_err(ds[n], dso[n]):.1e}" for n in range(M-1)])
print("per-interval dM err", [f"{rel_err(dm[n], dmo[n]):.1e}" for n in range(M-1)])
