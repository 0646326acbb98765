"""
Posterior reconstruction and the x wire format (rank 2).

CPU: the numpy oracle against golden vectors produced by the reference's own lambdified
polynomials (symbolics.get_local_polynomials driven as in example_500D.ipynb cell 29), and
the host-side index/pack helpers against the notebook's recipe.
GPU: csrc/posterior.cuh through the C ABI against the same goldens and the oracle.
Tolerance: 1e-10 relative (max norm), as for the free-energy path.
"""
from pathlib import Path

import numpy as np
import pytest

from helpers import rel_err

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["D5_irregular", "D1_single", "D12_uniform"]
RTOL = 1.0e-10


def _load(name):
    z = np.load(GOLDEN / f"posterior_{name}.npz")
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_polynomials(name):
    from oracle import posterior_oracle as O
    g = _load(name)
    tt, mt, st = O.reconstruct(g["mean_pts"], g["vars_pts"], float(g["t0"]), g["obs_times"],
                               float(g["tf"]), int(g["num"]))
    assert tt.shape == g["tt"].shape and mt.shape == g["mt"].shape
    assert np.array_equal(tt, g["tt"])
    assert rel_err(mt, g["mt"]) <= 1e-13 and rel_err(st, g["st"]) <= 1e-13
    # the polynomials interpolate: at the left knot of every interval m(t) is the support point
    num = int(g["num"])
    assert rel_err(mt[:, ::num][:, :-1], g["mean_pts"][:, 0:-3:3]) <= 1e-13


def test_support_indices_and_pack_follow_the_notebook():
    import meanfieldvargp_b200 as mf
    # example_500D.ipynb cell 17 (:356-384) on a small grid
    obs_idk, n_t = [25, 50, 75], 101
    mp_idk, sp_idk = mf.support_indices(obs_idk, n_t)
    M = len(obs_idk)
    assert mp_idk.size == 3 * M + 4 and sp_idk.size == 2 * M + 3
    time_space = [0, *obs_idk, n_t - 1]
    ref_m, ref_s = [], []
    for n in range(len(time_space) - 1):
        ref_m.extend(np.linspace(time_space[n], time_space[n + 1], num=4, endpoint=True, dtype=int))
        ref_s.extend(np.linspace(time_space[n], time_space[n + 1], num=3, endpoint=True, dtype=int))
    assert np.array_equal(mp_idk, np.unique(ref_m)) and np.array_equal(sp_idk, np.unique(ref_s))
    assert set(obs_idk) <= set(mp_idk) and set(obs_idk) <= set(sp_idk)
    rng = np.random.default_rng(0)
    D = 4
    mp, sp = rng.normal(size=(D, 3 * M + 4)), 4.0 + rng.random((D, 2 * M + 3))
    x = mf.pack(mp, sp)
    assert x.size == D * (5 * M + 7)
    assert np.array_equal(x[:mp.size], mp.ravel()) and np.allclose(np.exp(x[mp.size:]), sp.ravel())
    mp2, sp2 = mf.unpack(x, D, M)
    assert np.array_equal(mp2, mp) and rel_err(sp2, sp) <= 1e-15
    with pytest.raises(ValueError):
        mf.unpack(x[:-1], D, M)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_polynomials(name):
    import meanfieldvargp_b200 as mf
    g = _load(name)
    tt, mt, st = mf.reconstruct(g["mean_pts"], g["vars_pts"], float(g["t0"]), g["obs_times"],
                                float(g["tf"]), int(g["num"]))
    assert rel_err(tt, g["tt"]) <= 1e-15
    assert rel_err(mt, g["mt"]) <= RTOL and rel_err(st, g["st"]) <= RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("D,M,num", [(1, 1, 1), (3, 2, 5), (40, 19, 100), (500, 16, 100), (7, 300, 33)])
def test_gpu_matches_oracle_and_tensor_api(D, M, num):
    import torch
    import meanfieldvargp_b200 as mf
    from oracle import posterior_oracle as O
    rng = np.random.default_rng(D * 1000 + M)
    obs_times = np.cumsum(0.1 + rng.random(M))
    t0, tf = 0.0, float(obs_times[-1] + 0.37)
    mp = rng.normal(2.3, 3.6, size=(D, 3 * M + 4))
    sp = 4.0 + 2.0 * rng.random((D, 2 * M + 3))
    tt_o, mt_o, st_o = O.reconstruct(mp, sp, t0, obs_times, tf, num)
    tt, mt, st = mf.reconstruct(mp, sp, t0, obs_times, tf, num)
    assert mt.shape == (D, (M + 1) * num + 1)
    assert rel_err(tt, tt_o) <= 1e-15 and rel_err(mt, mt_o) <= RTOL and rel_err(st, st_o) <= RTOL
    # CUDA tensors in -> tensors out; log-variance block of x handed over directly
    x = mf.pack(mp, sp)
    xd = torch.tensor(x, dtype=torch.float64, device="cuda")
    mpd = xd[:mp.size].view(D, -1)
    lsd = xd[mp.size:].view(D, -1)
    tt2, mt2, st2 = mf.reconstruct(mpd, lsd, t0, obs_times, tf, num, log_vars=True)
    assert mt2.is_cuda and rel_err(mt2.cpu().numpy(), mt_o) <= RTOL
    assert rel_err(st2.cpu().numpy(), st_o) <= RTOL and rel_err(tt2.cpu().numpy(), tt_o) <= 1e-15


@pytest.mark.gpu
def test_reconstruct_after_find_minimum_interpolates_support_points():
    """End of the notebook flow: optimise, reconstruct, and the fine-grid curves pass through
    the optimised support points at the interval boundaries."""
    import contextlib
    import io
    import meanfieldvargp_b200 as mf
    from helpers import make_free_energy, synthetic_l96
    D, M = 40, 8
    g = synthetic_l96(D, M, seed=3)
    fe = make_free_energy(g)
    with contextlib.redirect_stdout(io.StringIO()):
        x, fx, _ = fe.find_minimum(g["x0"], max_iter=10)
    mp, sp = mf.unpack(x, D, M)
    t0, tf = 0.0, float(g["obs_times"][-1] + 0.25)
    tt, mt, st = mf.reconstruct(mp, sp, t0, g["obs_times"], tf, num=50)
    assert rel_err(mt[:, ::50][:, :-1], mp[:, 0:-3:3]) <= 1e-12
    assert rel_err(st[:, ::50][:, :-1], sp[:, 0:-2:2]) <= 1e-12
    assert np.all(np.diff(tt) > 0)


def test_initial_path_and_state_follow_the_notebook_cells():
    """example_500D.ipynb cells 14, 17, 19 written out, against mf.initial_path / initial_state"""
    from scipy.interpolate import splev, splrep
    from scipy.signal import savgol_filter
    import meanfieldvargp_b200 as mf
    rng = np.random.default_rng(3)
    D, dt, t0, tf = 6, 0.01, 0.0, 3.0
    tk = np.arange(t0, tf + dt, dt)
    path = np.cumsum(rng.standard_normal((tk.size, D)), axis=0) * 0.1
    h_mask = [True, False, True, True, False, True]
    obs_idk = np.array([50, 100, 150, 200, 250])
    obs_val = path[obs_idk][:, h_mask] + 0.1 * rng.standard_normal((obs_idk.size, sum(h_mask)))
    # cell 14
    spl_timex = np.array([t0, *(obs_idk * dt), tf])
    spl_value = np.array([obs_val[0], *obs_val, obs_val[-1]])
    want, j = [], 0
    for i in range(D):
        if h_mask[i]:
            want.append(splev(tk, splrep(spl_timex, spl_value[:, j], k=3)))
            j += 1
        else:
            want.append(savgol_filter(path[:, i], 51, 3))
    want = np.array(want)
    got = mf.initial_path(tk, obs_idk, obs_val, h_mask, path)
    assert got.shape == (D, tk.size) and np.allclose(got, want, rtol=0, atol=1e-12)
    # cells 17 + 19
    x, mp_idk, sp_idk = mf.initial_state(got, obs_idk, rng=np.random.default_rng(9))
    M = obs_idk.size
    assert len(mp_idk) == 3 * M + 4 and len(sp_idk) == 2 * M + 3
    r = np.random.default_rng(9)
    sp = np.log(4.0 * np.ones((D, len(sp_idk))) + 2.0 * r.random((D, len(sp_idk))))
    assert np.array_equal(x, np.concatenate((want[:, mp_idk].flatten(), sp.flatten())))
    with pytest.raises(ValueError, match="sample_path"):
        mf.initial_path(tk, obs_idk, obs_val, h_mask)


# ---- initial state x0 against the reference's notebook cells ------------
class _LegacyRand(object):
    """the notebook's unseeded global ``np.random.rand`` (example_500D.ipynb cell 19), seeded"""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def random(self, shape):
        return self.rs.random_sample(shape)


def _check_initial_state(g, sample_path, tk):
    import meanfieldvargp_b200 as mf
    # the path the notebook started from, regenerated bit for bit
    assert np.array_equal(sample_path[1000], g["path_row_1000"])
    assert np.array_equal(np.array([sample_path.sum(), np.abs(sample_path).sum()]), g["path_checksum"])
    h_mask = [bool(b) for b in g["h_mask"]]
    x0_path = mf.initial_path(tk, g["obs_idk"], g["obs_val"], h_mask, sample_path,
                              tf=float(g["tf"]), dt=float(g["dt"]))
    assert rel_err(x0_path[:, ::400], g["x0_path_probe"]) <= 1e-14       # cell 14
    x, mp_idk, sp_idk = mf.initial_state(x0_path, g["obs_idk"], rng=_LegacyRand(int(g["np_random_seed"])))
    assert np.array_equal(mp_idk, g["mp_idk"]) and np.array_equal(sp_idk, g["sp_idk"])   # cell 17
    D, M = int(g["D"]), len(g["obs_idk"])
    mp, sp = mf.unpack(x, D, M)
    assert rel_err(mp, g["mp_t0"]) <= 1e-14                                # cell 19
    assert rel_err(np.log(sp), g["sp_t0"]) <= 1e-14
    assert rel_err(x, g["x"]) <= 1e-14                                     # cell 23's argument


def test_initial_state_matches_the_reference_notebook_cells():
    """tests/golden/initial_L96_24p.npz: cells 8, 10, 14, 17, 19 of examples/example_500D.ipynb
    exec'ed from the notebook itself (oracle/ref_harness/make_golden_initial.py)"""
    from oracle import trajectory_oracle as O
    g = _load_any("initial_L96_24p")
    tk, path = O.make_trajectory(g["sigma"], float(g["theta"][0]), int(g["D"]), float(g["t0"]),
                                 float(g["tf"]), float(g["dt"]), r_seed=int(g["r_seed"]))
    _check_initial_state(g, path, tk)


@pytest.mark.gpu
def test_initial_state_from_the_device_sample_path():
    """the same, the sample path simulated on the device (csrc/trajectory.cuh, bit-identical)"""
    import meanfieldvargp_b200 as mf
    g = _load_any("initial_L96_24p")
    tk, x = mf.simulate_l96(g["sigma"], float(g["theta"][0]), int(g["D"]), float(g["t0"]),
                            float(g["tf"]), float(g["dt"]), r_seed=int(g["r_seed"]))
    _check_initial_state(g, x.cpu().numpy(), tk.cpu().numpy())


@pytest.mark.gpu
def test_initial_state_device_path_equals_host_path():
    """mfvgp_initial_state (gather + logs on the device) against the numpy helper, same draws"""
    import torch
    import meanfieldvargp_b200 as mf
    g = _load_any("initial_L96_24p")
    rng = np.random.default_rng(3)
    path = rng.normal(size=(int(g["D"]), 4002))
    xh, mi, si = mf.initial_state(path, g["obs_idk"], rng=_LegacyRand(5))
    xd, mi2, si2 = mf.initial_state(torch.tensor(path, device="cuda"), g["obs_idk"], rng=_LegacyRand(5))
    assert xd.is_cuda and np.array_equal(mi, mi2) and np.array_equal(si, si2)
    nm = path.shape[0] * len(mi)
    assert np.array_equal(xd.cpu().numpy()[:nm], xh[:nm])
    assert rel_err(xd.cpu().numpy()[nm:], xh[nm:]) <= 1e-15


def _load_any(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}
