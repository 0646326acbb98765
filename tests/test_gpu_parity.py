"""
GPU parity tests: the CUDA path (through the C ABI and the FreeEnergy drop-in)
against (a) the committed golden vectors produced by the unmodified reference
and (b) the numpy oracle on the same seeded inputs.

Tolerance: BASELINE.json north_star -- relative 1e-10 in FP64.
"""
import numpy as np
import pytest

from helpers import (EVAL_CASES, load_golden, make_free_energy, make_oracle, rel_err,
                     synthetic_l96)

pytestmark = pytest.mark.gpu
RTOL = 1.0e-10


@pytest.mark.parametrize("name", EVAL_CASES)
def test_ecost_matches_reference_golden(name):
    g = load_golden("eval", name)
    fe = make_free_energy(g)
    F, grad = fe.E_cost(g["x0"])
    assert abs(F - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad, g["grad"]) <= RTOL
    assert abs(fe.last_terms["E0"] - g["E0"]) <= RTOL * abs(g["F"])
    assert abs(fe.last_terms["Esde"] - g["Esde"]) <= RTOL * abs(g["F"])
    assert abs(fe.last_terms["Eobs"] - g["Eobs"]) <= RTOL * abs(g["F"])
    # free_energy.py:707 -- value-only call
    assert abs(fe.E_cost(g["x0"], output_gradients=False) - g["F_only"]) <= RTOL * abs(g["F"])
    assert fe.kernel_launches > 0


@pytest.mark.parametrize("name", EVAL_CASES)
def test_esde_matches_reference_golden(name):
    g = load_golden("eval", name)
    fe = make_free_energy(g)
    D, M = fe.dim_D, fe.num_M
    mp = g["x0"][:fe.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][fe.num_mp:].reshape(D, 2 * M + 3))
    Esde, dm, ds = fe.E_sde(mp, sp)
    assert dm.shape == (M - 1, D, 4) and ds.shape == (M - 1, D, 3)
    assert abs(Esde - g["Esde"]) <= RTOL * abs(g["Esde"])
    assert rel_err(dm, g["dEsde_dm"]) <= RTOL
    assert rel_err(ds, g["dEsde_ds"]) <= RTOL
    E_only, none_m, none_s = fe.E_sde(mp, sp, output_gradients=False)
    assert none_m is None and none_s is None
    assert abs(E_only - g["Esde"]) <= RTOL * abs(g["Esde"])


def test_deterministic_bitwise():
    g = load_golden("eval", "L96_40")
    fe = make_free_energy(g)
    F1, g1 = fe.E_cost(g["x0"])
    F2, g2 = fe.E_cost(g["x0"])
    assert F1 == F2 and np.array_equal(g1, g2)


@pytest.mark.parametrize("D,M", [(4, 2), (5, 3), (26, 4), (27, 5), (58, 3), (59, 6),
                                 (123, 4), (300, 9)])
def test_l96_matches_oracle_random_sizes(D, M):
    """Ragged sizes around the dim-tile edges (TE-6 = 26, 58, 122)."""
    g = synthetic_l96(D, M, seed=100 + D)
    fe = make_free_energy(g)
    F, grad = fe.E_cost(g["x0"])
    Fo, go = make_oracle(g, adaptive=False).E_cost(g["x0"])
    assert abs(F - Fo) <= RTOL * abs(Fo)
    assert rel_err(grad, go) <= RTOL


def test_tensor_api_zero_copy():
    import torch
    g = load_golden("eval", "L96_40")
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    F, grad = fe.E_cost(x)
    assert grad.is_cuda and grad.dtype == torch.float64
    assert abs(F.item() - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad.cpu().numpy(), g["grad"]) <= RTOL


def test_linearity_in_obs_noise_full_size():
    """
    Size-independent property at the north-star size (L96 D=500, M=16):
    Esde does not depend on the observation model, E_obs does.
    """
    g = synthetic_l96(500, 16, seed=5)
    fe1 = make_free_energy(g)
    F1, _ = fe1.E_cost(g["x0"])
    t1 = dict(fe1.last_terms)
    g2 = dict(g)
    g2["obs_values"] = g["obs_values"] + 1.0
    fe2 = make_free_energy(g2)
    fe2.E_cost(g["x0"])
    t2 = fe2.last_terms
    assert t1["Esde"] == t2["Esde"] and t1["E0"] == t2["E0"]
    assert t1["Eobs"] != t2["Eobs"]


def test_nonfinite_raises_like_reference():
    g = load_golden("eval", "L96_8")
    fe = make_free_energy(g)
    x = g["x0"].copy()
    x[3] = np.nan
    with pytest.raises(RuntimeError, match="is not a finite number"):
        fe.E_cost(x)


@pytest.mark.parametrize("name", ["DW", "L63", "L96_67p"])
def test_ekl0_and_eobs_standalone(name):
    """free_energy.py:281-336 and :567-649 as separate calls, vs the oracle."""
    g = load_golden("eval", name)
    fe = make_free_energy(g)
    o = make_oracle(g)
    D, M = fe.dim_D, fe.num_M
    mp = g["x0"][:fe.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][fe.num_mp:].reshape(D, 2 * M + 3))
    E0, dm0, ds0 = fe.E_kl0(mp[:, 0], sp[:, 0])
    E0o, dm0o, ds0o = o.E_kl0(mp[:, 0], sp[:, 0])
    assert abs(E0 - g["E0"]) <= RTOL * abs(g["F"]) and abs(E0 - E0o) <= RTOL * abs(g["F"])
    assert rel_err(dm0, dm0o) <= RTOL and rel_err(ds0, ds0o) <= RTOL
    assert fe.E_kl0(mp[:, 0], sp[:, 0], output_gradients=False)[1] is None
    Eobs, dm, ds = fe.E_obs(mp[o.ix_ikm], sp[o.ix_iks])
    Eobso, dmo, dso = o.E_obs(mp[o.ix_ikm], sp[o.ix_iks])
    assert abs(Eobs - g["Eobs"]) <= RTOL * abs(g["F"])
    assert rel_err(np.atleast_2d(dm), dmo) <= RTOL and rel_err(np.atleast_2d(ds), dso) <= RTOL
    if fe.dim_d == 1:
        assert dm.ndim == 1           # squeezed like the reference (:637-642)


def test_split_rule_equals_42_node_rule(monkeypatch):
    """
    The default L96 kernel integrates the polynomial part of the integrands on 7
    Gauss-Legendre nodes and the rational part on scipy's 42 Kronrod nodes; the
    node-for-node 42-point kernel (MFVGP_RULE=gk42) must give the same numbers.
    """
    for name in ("L96_40", "L96_67p", "L96_5_irregular"):
        g = load_golden("eval", name)
        monkeypatch.setenv("MFVGP_RULE", "gk42")
        fe42 = make_free_energy(g)
        F42, g42 = fe42.E_cost(g["x0"])
        monkeypatch.delenv("MFVGP_RULE")
        fes = make_free_energy(g)
        Fs, gs = fes.E_cost(g["x0"])
        assert abs(Fs - F42) <= 1e-13 * abs(F42)
        assert rel_err(gs, g42) <= 1e-12
        assert abs(F42 - g["F"]) <= RTOL * abs(g["F"]) and rel_err(g42, g["grad"]) <= RTOL


def _roughen(g, D, M, rough):
    x = g["x0"].copy()
    if rough:
        rng = np.random.default_rng(M)
        sp = x[D * (3 * M + 4):].reshape(D, 2 * M + 3)
        cols = rng.choice(2 * M + 3, size=max(2, M // 6), replace=False)
        sp[:, cols] = np.log(10.0 ** rng.uniform(-1.7, -0.7, size=(D, cols.size)))
    return x


def _check_path_equals_split(path, D, M, rough, monkeypatch, ftile=None):
    g = synthetic_l96(D, M, seed=D + M)
    x = _roughen(g, D, M, rough)
    monkeypatch.setenv("MFVGP_PATH", "split")
    monkeypatch.setenv("MFVGP_COOP", "0")
    fe0 = make_free_energy(g)
    F0, g0 = fe0.E_cost(x)
    info0 = fe0.refine_info()
    if path == "coop":              # split path with the cooperative small-problem kernels
        monkeypatch.setenv("MFVGP_COOP", "1")
    else:
        monkeypatch.setenv("MFVGP_PATH", path)
    if ftile:
        monkeypatch.setenv("MFVGP_FTILE", ftile)
    fe1 = make_free_energy(g)
    monkeypatch.delenv("MFVGP_PATH")
    monkeypatch.delenv("MFVGP_COOP")
    monkeypatch.delenv("MFVGP_FTILE", raising=False)
    F1, g1 = fe1.E_cost(x)
    assert np.array_equal(fe1.refine_info(), info0)
    if rough:
        assert info0.max() > 0
    assert abs(F1 - F0) <= 1e-13 * abs(F0)
    assert rel_err(g1, g0) <= 1e-12
    assert abs(fe1.E_cost(x, output_gradients=False) - F0) <= 1e-13 * abs(F0)
    if D * (M - 1) < 20000 and info0.max() < 90:      # limit=100 hit => not comparable
        Fo, go = make_oracle(g, adaptive=True).E_cost(x)
        assert abs(F1 - Fo) <= RTOL * abs(Fo) and rel_err(g1, go) <= RTOL


@pytest.mark.parametrize("D,M,rough,ftile", [
    (4, 2, False, None), (5, 3, False, "32,1"), (26, 9, False, "32,3"), (27, 5, True, None),
    (40, 7, False, "64,2"), (67, 12, True, "64,5"), (123, 40, True, "128,7"),
    (300, 300, False, None), (300, 290, True, "128,64"), (64, 1500, True, None),
    (500, 16, False, "128,1"), (1000, 130, True, "128,16")])
def test_walk_ecost_equals_split(D, M, rough, ftile, monkeypatch):
    """
    The walk kernel (MFVGP_PATH=walk, l96_walk.cuh: a CTA walks along a run of
    intervals, moment form of the rational part, gradient written directly) against
    the split kernel + assemble_kernel; run lengths chosen to hit ragged last runs,
    single-interval runs and the maximum run length.
    """
    _check_path_equals_split("walk", D, M, rough, monkeypatch, ftile)


@pytest.mark.parametrize("D,M,rough", [(4, 2, False), (5, 3, False), (26, 9, True), (27, 5, True),
                                       (40, 7, False), (67, 12, True), (123, 40, True),
                                       (500, 16, False), (500, 16, True), (300, 200, True)])
def test_coop_ecost_equals_split(D, M, rough, monkeypatch):
    """
    Small problems run esde_l96_coop_kernel (8 lanes per cell, guard folded in by the
    last CTA of each interval) + assemble_tail_kernel (assembly + final sums); they must
    reproduce the one-thread-per-cell split kernel + assemble_kernel + final_kernel.
    """
    _check_path_equals_split("coop", D, M, rough, monkeypatch)
    # E_sde alone goes through the same kernel
    g = synthetic_l96(D, M, seed=D + M)
    fe = make_free_energy(g)
    mp = g["x0"][:fe.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][fe.num_mp:].reshape(D, 2 * M + 3))
    Es, dm, ds = fe.E_sde(mp, sp)
    monkeypatch.setenv("MFVGP_COOP", "0")
    fe0 = make_free_energy(g)
    monkeypatch.delenv("MFVGP_COOP")
    Es0, dm0, ds0 = fe0.E_sde(mp, sp)
    assert abs(Es - Es0) <= 1e-13 * abs(Es0)
    assert rel_err(dm, dm0) <= 1e-12 and rel_err(ds, ds0) <= 1e-12


@pytest.mark.parametrize("name", EVAL_CASES)
def test_walk_matches_reference_golden(name, monkeypatch):
    g = load_golden("eval", name)
    if g["system"] != "L96":
        pytest.skip("walk kernel is the Lorenz 96 path")
    monkeypatch.setenv("MFVGP_PATH", "walk")
    fe = make_free_energy(g)
    monkeypatch.delenv("MFVGP_PATH")
    F, grad = fe.E_cost(g["x0"])
    assert abs(F - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad, g["grad"]) <= RTOL
    for k in ("E0", "Esde", "Eobs"):
        assert abs(fe.last_terms[k] - g[k]) <= RTOL * abs(g["F"])


def test_integration_stub_of_the_docs():
    """
    INTEGRATION.md section 3: the few lines of ctypes a maintainer of the reference would add
    (mfvgp_create / set_* / mfvgp_ecost_host on plain numpy buffers) -- exercised verbatim.
    """
    import ctypes
    from meanfieldvargp_b200 import _lib
    g = load_golden("eval", "L96_40")
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    P = ctypes.c_void_p
    D, M = g["sigma"].size, len(g["obs_times"])
    h = P()
    assert lib.mfvgp_create(ctypes.byref(h), 3, D, M, D, 0, 0, -1) == 0
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)          # noqa: E731
    s, t = f64(g["sigma"]), f64(g["theta"])
    assert lib.mfvgp_set_sde(h, P(s.ctypes.data), P(t.ctypes.data), t.size) == 0
    mu0, tau0 = f64(np.broadcast_to(g["mu0"], (D,))), f64(np.broadcast_to(g["tau0"], (D,)))
    assert lib.mfvgp_set_prior(h, P(mu0.ctypes.data), P(tau0.ctypes.data)) == 0
    ot, ov = f64(g["obs_times"]), f64(g["obs_values"]).reshape(M, D)
    on = f64(np.broadcast_to(g["obs_noise"], (D,)))
    assert lib.mfvgp_set_obs(h, P(ot.ctypes.data), P(ov.ctypes.data), P(on.ctypes.data), None) == 0
    x = f64(g["x0"])
    out, grad, st = np.empty(4), np.empty_like(x), ctypes.c_int32()
    rc = lib.mfvgp_ecost_host(h, P(x.ctypes.data), 1, P(out.ctypes.data), P(grad.ctypes.data),
                              ctypes.byref(st))
    assert rc == 0 and (st.value & 7) == 0
    assert abs(out[0] - g["F"]) <= RTOL * abs(g["F"]) and rel_err(grad, g["grad"]) <= RTOL
    # value only: the gradient buffer may be NULL
    rc = lib.mfvgp_ecost_host(h, P(x.ctypes.data), 0, P(out.ctypes.data), None, ctypes.byref(st))
    assert rc == 0 and abs(out[0] - g["F"]) <= RTOL * abs(g["F"])
    # calls in the wrong order fail with a message instead of computing garbage
    h2 = P()
    assert lib.mfvgp_create(ctypes.byref(h2), 3, D, M, D, 0, 0, -1) == 0
    assert lib.mfvgp_ecost_host(h2, P(x.ctypes.data), 0, P(out.ctypes.data), None,
                                ctypes.byref(st)) == -3
    lib.mfvgp_last_error.restype = ctypes.c_char_p
    assert b"not called" in lib.mfvgp_last_error()
    assert lib.mfvgp_destroy(h) == 0 and lib.mfvgp_destroy(h2) == 0


def test_single_observed_dimension_and_scalar_noise():
    """d = 1: obs_values may come as a 1-D array and obs_noise as a scalar (free_energy.py:136-147)."""
    g = synthetic_l96(12, 5, seed=4)
    mask = [False] * 12
    mask[7] = True
    g["obs_mask"] = mask
    g["obs_values"] = g["obs_values"][:, 0].copy()          # shape (M,)
    g["obs_noise"] = np.array(1.7)
    fe = make_free_energy(g)
    F, grad = fe.E_cost(g["x0"])
    g2 = dict(g)
    g2["obs_values"] = g["obs_values"].reshape(-1, 1)
    g2["obs_noise"] = np.array([1.7])
    Fo, go = make_oracle(g2, adaptive=True).E_cost(g["x0"])
    assert abs(F - Fo) <= RTOL * abs(Fo) and rel_err(grad, go) <= RTOL


def test_scg_class_surface_and_tensor_ekl0_eobs(capsys):
    """SCG(func, options)(x0) -> (x, fx), .statistics (scaled_cg.py:25, 93-105, 388); the
    ix_ikm / ix_iks attributes (free_energy.py:176-177); E_kl0 / E_obs on CUDA tensors"""
    import torch
    import meanfieldvargp_b200 as mf
    g = load_golden("scg", "L96_8")
    fe = make_free_energy(g)
    opt = mf.SCG(fe.E_cost, {"max_it": int(g["max_iter"]), "x_tol": 1.0e-5, "f_tol": 1.0e-5})
    with pytest.raises(NotImplementedError):
        opt.statistics
    x, fx = opt(g["x0"].copy())
    st = opt.statistics
    assert st["nit"] == int(g["nit"]) and st["func_eval"] == int(g["func_eval"])
    assert abs(fx - float(g["fx_opt"])) <= 1e-8 * abs(float(g["fx_opt"]))
    assert rel_err(x, g["x_opt"]) <= 1e-6
    assert "Maximum number of iterations" in capsys.readouterr().out
    with pytest.raises(TypeError):
        mf.SCG(lambda v: (0.0, v))(g["x0"])
    # E_obs through the reference's own index grids, numpy and tensor paths
    D, M = fe.dim_D, fe.num_M
    mp = g["x0"][:fe.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][fe.num_mp:].reshape(D, 2 * M + 3))
    o = make_oracle(g)
    Eo, dmo, dso = o.E_obs(mp[o.ix_ikm], sp[o.ix_iks])
    E1, dm1, ds1 = fe.E_obs(mp[fe.ix_ikm], sp[fe.ix_iks])
    E2, dm2, ds2 = fe.E_obs(torch.tensor(mp[fe.ix_ikm], device="cuda"),
                            torch.tensor(sp[fe.ix_iks], device="cuda"))
    assert dm2.is_cuda and dm2.shape == dmo.shape
    for E, dm, ds in ((E1, dm1, ds1), (float(E2), dm2.cpu().numpy(), ds2.cpu().numpy())):
        assert abs(E - Eo) <= 1e-12 * abs(Eo) and rel_err(dm, dmo) <= 1e-12 and rel_err(ds, dso) <= 1e-12
    K0, a0, b0 = o.E_kl0(mp[:, 0], sp[:, 0])
    K2, a2, b2 = fe.E_kl0(torch.tensor(mp[:, 0], device="cuda"), torch.tensor(sp[:, 0], device="cuda"))
    assert a2.is_cuda and abs(float(K2) - K0) <= 1e-12 * abs(K0)
    assert rel_err(a2.cpu().numpy(), a0) <= 1e-12 and rel_err(b2.cpu().numpy(), b0) <= 1e-12
    assert fe.E_kl0(torch.tensor(mp[:, 0], device="cuda"), torch.tensor(sp[:, 0], device="cuda"),
                    output_gradients=False)[1] is None
