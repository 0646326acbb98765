"""
Generates tests/golden/posterior_*.npz by running the UNMODIFIED reference:
``symbolics.get_local_polynomials()`` (src/numerical/symbolics.py:102-149) driven exactly
as examples/example_500D.ipynb cell 29 (:537-575) drives it.  Run in the build container
(needs /root/reference and sympy):

    python oracle/ref_harness/make_golden_posterior.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref_harness import harness  # noqa: E402


def reference_reconstruct(mp_tf, sp_tf, t0, obs_times, tf, num):
    """cell 29, verbatim control flow (the triple loop), on the reference's lambdas."""
    harness.install()
    from src.numerical.symbolics import get_local_polynomials
    poly_mean, poly_vars = get_local_polynomials()
    D = mp_tf.shape[0]
    tt, mt, st = [], [], []
    Tx = [t0, *obs_times, tf]
    for n in range(0, len(Tx) - 1):
        ti, tj = Tx[n], Tx[n + 1]
        delta_t = np.abs(tj - ti)
        h = float(delta_t / 3.0)
        c = float(delta_t / 2.0)
        nth_mean_points = mp_tf[:, (3 * n): (3 * n) + 4]
        nth_vars_points = sp_tf[:, (2 * n): (2 * n) + 3]
        for t in np.linspace(ti, tj, num=num, endpoint=False):
            mt_, st_ = [], []
            for i in range(D):
                par_m = [ti, ti + h, ti + (2 * h), ti + (3 * h), *nth_mean_points[i]]
                par_s = [ti, ti + c, ti + (2 * c), *nth_vars_points[i]]
                mt_.append(poly_mean(t, *par_m))
                st_.append(poly_vars(t, *par_s))
            mt.append(mt_)
            st.append(st_)
            tt.append(t)
    mt_, st_ = [], []
    for i in range(D):
        par_m = [ti, ti + h, ti + (2 * h), ti + (3 * h), *nth_mean_points[i]]
        par_s = [ti, ti + c, ti + (2 * c), *nth_vars_points[i]]
        mt_.append(poly_mean(tf, *par_m))
        st_.append(poly_vars(tf, *par_s))
    mt.append(mt_)
    st.append(st_)
    tt.append(tf)
    return np.array(tt), np.array(mt).T, np.array(st).T


CASES = {
    # name: (D, obs_times, t0, tf, num, seed)
    "D5_irregular": (5, [0.3, 0.5, 1.1, 1.25, 2.0], 0.0, 2.4, 7, 11),
    "D1_single": (1, [0.5], 0.0, 1.0, 100, 12),
    "D12_uniform": (12, list(0.25 * np.arange(1, 9)), 0.0, 2.25, 100, 13),
}


def main():
    out = ROOT / "tests" / "golden"
    for name, (D, obs_times, t0, tf, num, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        M = len(obs_times)
        mp = rng.normal(2.3, 3.6, size=(D, 3 * M + 4))
        sp = 4.0 + 2.0 * rng.random((D, 2 * M + 3))
        tt, mt, st = reference_reconstruct(mp, sp, t0, obs_times, tf, num)
        np.savez_compressed(out / f"posterior_{name}.npz", mean_pts=mp, vars_pts=sp,
                            obs_times=np.array(obs_times, dtype=float), t0=t0, tf=tf, num=num,
                            tt=tt, mt=mt, st=st)
        print(name, tt.shape, mt.shape, st.shape)


if __name__ == "__main__":
    main()
