"""
TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's posterior
reconstruction, the checker for csrc/posterior.cuh.  Only tests/, smoke() and bench.py's
CPU legs may import this.

Follows examples/example_500D.ipynb cell 29 (:537-575): for every interval of
Tx = [t0, *obs_times, tf] the cubic / quadratic Lagrange polynomials of
src/numerical/symbolics.py:9-100 (LagrangePolynomial) on the knots ti + k h, h = dt/3
(mean) and ti + k c, c = dt/2 (variance), sampled at np.linspace(ti, tj, num,
endpoint=False); the final sample is tf on the last interval's polynomial.

Pinned against the reference's own lambdified polynomials
(symbolics.get_local_polynomials, :102-149) by tests/golden/posterior_*.npz
(oracle/ref_harness/make_golden_posterior.py).
"""
import numpy as np


def lagrange(t, knots, values):
    """sum_k values[..., k] prod_{l != k} (t - knots[l]) / (knots[k] - knots[l])  (symbolics.py:64-90)."""
    t = np.asarray(t, dtype=float)
    out = 0.0
    K = len(knots)
    for k in range(K):
        w = np.ones_like(t)
        for l in range(K):
            if l != k:
                w = w * ((t - knots[l]) / (knots[k] - knots[l]))
        out = out + values[..., k, None] * w
    return out


def reconstruct(mean_pts, vars_pts, t0, obs_times, tf, num=100):
    """(tt [P], mt [D, P], st [D, P]), P = (M+1) num + 1 -- cell 29 vectorised over D and t."""
    mean_pts = np.atleast_2d(np.asarray(mean_pts, dtype=float))
    vars_pts = np.atleast_2d(np.asarray(vars_pts, dtype=float))
    Tx = np.concatenate(([float(t0)], np.asarray(obs_times, dtype=float).ravel(), [float(tf)]))
    tt, mt, st = [], [], []
    for n in range(len(Tx) - 1):
        ti, tj = Tx[n], Tx[n + 1]
        delta_t = np.abs(tj - ti)
        h, c = float(delta_t / 3.0), float(delta_t / 2.0)
        mp = mean_pts[:, 3 * n:3 * n + 4]
        sp = vars_pts[:, 2 * n:2 * n + 3]
        kh = [ti, ti + h, ti + (2 * h), ti + (3 * h)]
        kc = [ti, ti + c, ti + (2 * c)]
        t = np.linspace(ti, tj, num=num, endpoint=False)
        tt.append(t)
        mt.append(lagrange(t, kh, mp))
        st.append(lagrange(t, kc, sp))
    # final point tf, on the last interval's polynomial (:561-572)
    tt.append(np.array([float(tf)]))
    mt.append(lagrange(np.array([float(tf)]), kh, mp))
    st.append(lagrange(np.array([float(tf)]), kc, sp))
    return np.concatenate(tt), np.concatenate(mt, axis=1), np.concatenate(st, axis=1)
