"""
Randomised parity sweep: ragged sizes, irregular observation times, per-dimension noise
levels, random observation masks and forcing -- every Lorenz 96 path (cooperative
small-problem kernels, walk kernel with random run lengths, split kernel) against the numpy
oracle with scipy's adaptive quad_vec.  Tolerance 1e-10 relative (BASELINE.json).
"""
import numpy as np
import pytest

from helpers import make_free_energy, make_oracle, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1.0e-10


def _random_problem(seed):
    rng = np.random.default_rng(seed)
    D = int(rng.integers(4, 141))
    M = int(rng.integers(2, 13))
    Nm, Ns = 3 * M + 4, 2 * M + 3
    obs_times = np.cumsum(rng.uniform(0.05, 0.6, size=M))
    mask = rng.random(D) < rng.uniform(0.2, 1.0)
    mask[rng.integers(0, D)] = True
    d = int(mask.sum())
    mp = rng.normal(2.0, 3.0, size=(D, Nm))
    sp = np.log(rng.uniform(0.5, 6.0, size=(D, Ns)))
    if seed % 3 == 0:                       # a few steep variance triples: adaptive refinement
        sp[rng.integers(0, D, size=3), rng.integers(0, Ns, size=3)] = np.log(0.12)
    return dict(system="L96", sigma=rng.uniform(5.0, 60.0, size=D),
                theta=np.array([rng.uniform(4.0, 10.0)]), mu0=rng.normal(5.0, 3.0, size=D),
                tau0=rng.uniform(1.0, 6.0, size=D), obs_times=obs_times,
                obs_values=rng.normal(2.0, 4.0, size=(M, d)), obs_noise=rng.uniform(0.5, 3.0, size=d),
                obs_mask=[bool(b) for b in mask], x0=np.concatenate((mp.ravel(), sp.ravel())))


@pytest.mark.parametrize("seed", range(24))
def test_random_problem_all_paths(seed, monkeypatch):
    g = _random_problem(seed)
    o = make_oracle(g, adaptive=True)
    Fo, go = o.E_cost(g["x0"])
    if max(max(n) for n in o.neval) >= 21 + 42 * 90:       # quad_vec hit limit=100: not comparable
        pytest.skip("reference integration did not converge")
    D, M = len(g["obs_mask"]), len(g["obs_times"])
    tn = 1 + seed % 5
    for env in ({}, {"MFVGP_PATH": "walk", "MFVGP_FTILE": f"{(32, 64, 128)[seed % 3]},{tn}"},
                {"MFVGP_PATH": "split", "MFVGP_COOP": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        fe = make_free_energy(g)
        for k in env:
            monkeypatch.delenv(k)
        F, grad = fe.E_cost(g["x0"])
        info = fe.refine_info()            # of the call above (a value-only call skips dS)
        assert abs(F - Fo) <= RTOL * abs(Fo), (env, F, Fo)
        assert rel_err(grad, go) <= RTOL, (env, rel_err(grad, go))
        assert abs(fe.E_cost(g["x0"], output_gradients=False) - Fo) <= RTOL * abs(Fo)
        # identical quad_vec decisions: k bisections <-> neval = 21 + 42 k (k = 0, 1: 63)
        nev = np.array(o.neval)
        assert np.array_equal(np.where(info[:, 0] > 0, 21 + 42 * info[:, 0], 63), nev[:, 0])
        assert np.array_equal(np.where(info[:, 1] > 0, 21 + 42 * info[:, 1], 63), nev[:, 2])
