"""
Device SCG (mfvgp_scg) against the reference's SCG.minimize trajectories
(golden vectors from the unmodified reference, scaled_cg.py:107-386).

SCG is a branchy iteration: the bar is the identical branch sequence (same
nit and func_eval, same accept/reject pattern visible in fx) with the energy
history matching to 1e-8 relative and the final iterate to 1e-6.
"""
import numpy as np
import pytest

from helpers import (SCG_BRANCH_CASES, SCG_CASES, load_golden, make_free_energy, make_oracle,
                     rel_err)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", SCG_CASES)
def test_scg_matches_reference_trajectory(name, capsys):
    g = load_golden("scg", name)
    fe = make_free_energy(g)
    x, fx, stats = fe.find_minimum(g["x0"].copy(), max_iter=int(g["max_iter"]))
    out = capsys.readouterr().out
    assert "Elapsed time [" in out and "Done!" in out
    assert stats["nit"] == int(g["nit"])
    assert stats["func_eval"] == int(g["func_eval"])
    nit = stats["nit"]
    assert rel_err(stats["fx"][:nit], g["fx_hist"][:nit]) <= 1.0e-8
    assert rel_err(stats["dfx"][:nit], g["dfx_hist"][:nit]) <= 1.0e-5
    assert abs(fx - float(g["fx_opt"])) <= 1.0e-8 * abs(float(g["fx_opt"]))
    assert rel_err(x, g["x_opt"]) <= 1.0e-6


VARIANTS = [{}, {"MFVGP_SCG_PIPELINE": "0"}, {"MFVGP_SCG_PIPELINE": "2"},
            {"MFVGP_SCG_CLUSTER": "0"}, {"MFVGP_SCG_FUSED": "0"},
            {"MFVGP_SCG_PIPELINE": "2", "MFVGP_SCG_CLUSTER": "0"}]


@pytest.mark.parametrize("name", SCG_BRANCH_CASES)
@pytest.mark.parametrize("variant", range(len(VARIANTS)))
def test_scg_rare_branches_match_reference(name, variant, monkeypatch):
    """
    Runs of the UNMODIFIED reference that reject trial points (scaled_cg.py:254-279: g_now <-
    grad_old, dfx = sum|grad_old|), find mu >= 0 and reset the direction (:194-197), find
    delta <= 0 and rescale (:233-236; the hilltop cases: a double-well mean path started on the
    unstable point between the wells) and restart after dim_x successes (:361-364): same nit,
    func_eval and fx / dfx history on the device, in
    every variant of the control loop -- chains enqueued ahead (and voided by a rejection),
    step by step, cluster and grid reductions, host-read records (MFVGP_SCG_FUSED=0).
    """
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    g = load_golden("scg", name)
    assert int(g["n_reject"]) >= 2                       # + mu >= 0, delta <= 0, restarts: see
    fe = make_free_energy(g)                             # test_reference_goldens_cover_every_rare_scg_branch
    x, fx, stats = fe.find_minimum(g["x0"].copy(), max_iter=int(g["max_iter"]))
    nit = int(g["nit"])
    assert stats["nit"] == nit and stats["func_eval"] == int(g["func_eval"])
    assert rel_err(stats["fx"][:nit], g["fx_hist"][:nit]) <= 1.0e-8
    assert rel_err(stats["dfx"][:nit], g["dfx_hist"][:nit]) <= 1.0e-6
    # the accept / reject pattern itself: fx repeats exactly where a trial point was rejected
    assert np.array_equal(np.diff(stats["fx"][:nit]) == 0.0, np.diff(g["fx_hist"][:nit]) == 0.0)
    assert abs(fx - float(g["fx_opt"])) <= 1.0e-8 * abs(float(g["fx_opt"]))
    assert rel_err(x, g["x_opt"]) <= 1.0e-6


WILD = {"OU2": ("OU", 2, 290291082)}


@pytest.mark.parametrize("case", sorted(WILD))
def test_scg_wild_start_raises_like_the_reference(case):
    """
    Starts with scattered log-variances (x0 + as N(0,1), as ~ 1): the interpolated variance dips
    to zero inside an interval, and the UNMODIFIED reference stops its SCG run with
    ``RuntimeError: ... is not a finite number`` (free_energy.py:313-316, 547-550) -- recorded
    while searching for starts that take SCG's rare branches (tools/scg_branch_search.py).  The
    device run must end the same way: the reference's exception, not a silent result.
    """
    name, M_, seed = WILD[case]
    g0 = load_golden("scg", name)
    D_ = int(np.atleast_1d(g0["sigma"]).size)
    M0 = int(g0["obs_times"].size)
    mp = g0["x0"][:D_ * (3 * M0 + 4)].reshape(D_, 3 * M0 + 4)[:, :3 * M_ + 4]
    sp = g0["x0"][D_ * (3 * M0 + 4):].reshape(D_, 2 * M0 + 3)[:, :2 * M_ + 3]
    g = dict(g0)
    g["obs_times"], g["obs_values"] = g0["obs_times"][:M_], g0["obs_values"][:M_]
    r = np.random.default_rng(seed)
    am, asg = float(r.choice([0.2, 0.5, 1.0, 2.0])), float(r.choice([0.6, 0.8, 1.0, 1.2]))
    nm = D_ * (3 * M_ + 4)
    x0 = np.concatenate((mp.ravel(), sp.ravel()))
    x0[:nm] += am * r.standard_normal(nm)
    x0[nm:] += asg * r.standard_normal(x0.size - nm)
    fe = make_free_energy(g)
    with pytest.raises(RuntimeError, match="is not a finite number"):
        fe.find_minimum(x0.copy(), max_iter=40)


def test_scg_tensor_in_place_and_floor():
    import torch
    g = load_golden("scg", "L96_8")
    fe = make_free_energy(g)
    x0 = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    # max_iter is floored at 10 (free_energy.py:815)
    x, fx, stats = fe.find_minimum(x0, max_iter=3)
    assert stats["fx"].size == 10 and x.is_cuda
    assert fx <= float(fe.E_cost(x0)[0].item())
    assert torch.equal(x0, torch.tensor(g["x0"], dtype=torch.float64, device="cuda"))


def test_scg_decreases_energy_full_size():
    """North-star size: D=500 synthetic; SCG must reduce F monotonically in fx."""
    from helpers import synthetic_l96
    g = synthetic_l96(500, 16, seed=11)
    fe = make_free_energy(g)
    F0 = fe.E_cost(g["x0"], output_gradients=False)
    x, fx, stats = fe.find_minimum(g["x0"], max_iter=12)
    assert fx < F0
    hist = stats["fx"][:stats["nit"]]
    assert np.all(np.diff(hist) <= 1e-9 * abs(F0))


# ---- pipelined chains, cluster kernels, guard in the tail kernel ---------------------------------
ROUGH = {"rough_L63": ("L63_T5", 3, 2.5, 4), "rough_L96": ("L96_8", 8, 2.0, 1),
         "rough_DW": ("DW", 1, 1.0, 3)}


def _rough(case):
    """Golden problems started from scattered log-variances (x0 + base N(0,1)): the quadratic
    model of the sigma step is far off there and SCG REJECTS most trial points (Delta < 0,
    scaled_cg.py:254-279).  The reference goldens never do (func_eval = 3 nit + 1 in all of them)
    and sane starts never did in a search either; these starts are numerically wild (interpolated
    variances near zero), so they pin the control flow -- rejected trial points void the chain
    enqueued ahead, the step-by-step path takes over -- not values against the oracle."""
    name, dims, base, seed = ROUGH[case]
    g = dict(load_golden("scg", name))
    n = g["x0"].size
    nm = dims * (3 * int(g["obs_times"].size) + 4)
    x0 = g["x0"].copy()
    x0[nm:] = x0[nm:] + base * np.random.default_rng(seed).normal(0.0, 1.0, n - nm)
    g["x0"] = x0
    return g


@pytest.mark.parametrize("case", ["L96_8", "L96_40", "rough_L63", "rough_L96", "rough_DW"])
def test_scg_switches_agree(case, monkeypatch):
    """pipelining on/off is bit-identical; cluster and grid reductions agree to rounding and take
    the same branches"""
    from helpers import synthetic_l96
    rough = case.startswith("rough")
    if rough:
        g, iters = _rough(case), 40
    elif case == "L96_40":
        g, iters = synthetic_l96(40, 20, seed=5), 60
    else:
        g, iters = load_golden("scg", case), 15
    res = {}
    for pipe in ("0", "1", "2"):                  # 2: a chain is enqueued ahead after EVERY chain
        for cl in ("0", "1"):
            monkeypatch.setenv("MFVGP_SCG_PIPELINE", pipe)
            monkeypatch.setenv("MFVGP_SCG_CLUSTER", cl)
            fe = make_free_energy(g)
            x, fx, st = fe.find_minimum(g["x0"].copy(), max_iter=iters)
            res[pipe, cl] = (x, fx, st["nit"], st["func_eval"], st["fx"].copy())
    for cl in ("0", "1"):
        for pipe in ("1", "2"):
            a, b = res["0", cl], res[pipe, cl]
            assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2:4] == b[2:4]
    if rough:                                     # there ARE rejected iterations
        assert res["1", "1"][3] < 3 * res["1", "1"][2] + 1
        return
    a, b = res["1", "0"], res["1", "1"]
    assert a[2:4] == b[2:4]
    assert rel_err(a[4], b[4]) <= 1.0e-10 and rel_err(a[0], b[0]) <= 1.0e-8


def test_scg_converges_with_a_chain_in_flight():
    """loose tolerances: the run stops (scaled_cg.py:306) while the next chain is already
    enqueued; that chain must be void and the result the accepted point"""
    g = load_golden("scg", "L96_8")
    xo, fxo, so = make_oracle(g, adaptive=False).find_minimum(g["x0"].copy(), max_iter=200,
                                                              x_tol=5.0e-2, f_tol=5.0e-2)
    assert so["nit"] < 200
    fe = make_free_energy(g)
    x, fx, st = fe.find_minimum(g["x0"].copy(), max_iter=200, x_tol=5.0e-2, f_tol=5.0e-2)
    assert st["nit"] == so["nit"] and st["func_eval"] == so["func_eval"]
    assert abs(fx - fxo) <= 1.0e-8 * abs(fxo) and rel_err(x, xo) <= 1.0e-6
    F, _ = fe.E_cost(x)
    assert abs(F - fx) <= 1.0e-12 * abs(fx)


@pytest.mark.parametrize("name", ["L96_8", "L96_40", "L96_6_refine", "L96_500p"])
def test_guard_in_tail_kernel_equals_guard_in_main_kernel(name, monkeypatch):
    """optimistic evaluations (host buffers) fold the tile partials and apply the guard in CTAs of
    the tail kernel instead of the main kernel's last CTAs: same bits, same flags"""
    kind = "scg" if name == "L96_8" else "eval"
    g = load_golden(kind, name)
    x0 = g["x0"] if "x0" in g else g["x"]
    out = {}
    for v in ("0", "1"):
        monkeypatch.setenv("MFVGP_GUARD_IN_TAIL", v)
        fe = make_free_energy(g)
        F, grad = fe.E_cost(np.array(x0, dtype=float))
        out[v] = (F, grad.copy())
    assert out["0"][0] == out["1"][0] and np.array_equal(out["0"][1], out["1"][1])


def test_scg_fallback_cluster_shape_agrees():
    """8 CTAs x 1024 threads (devices without 16-CTA clusters) against the default 16 x 512: the
    choice is cached per process, so the fallback runs in a subprocess"""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    from helpers import synthetic_l96
    code = (
        "import sys, json, contextlib, io; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "from helpers import make_free_energy, synthetic_l96\n"
        "g = synthetic_l96(40, 20, seed=5); fe = make_free_energy(g)\n"
        "with contextlib.redirect_stdout(io.StringIO()):\n"
        "    x, fx, st = fe.find_minimum(g['x0'].copy(), max_iter=60)\n"
        "print(json.dumps([int(st['nit']), int(st['func_eval']), float(fx)]))\n")
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, MFVGP_SCG_CLUSTER16="0"), timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    nit, nev, fx8 = json.loads(out.stdout.strip().splitlines()[-1])
    g = synthetic_l96(40, 20, seed=5)
    x, fx, st = make_free_energy(g).find_minimum(g["x0"].copy(), max_iter=60)
    assert (nit, nev) == (st["nit"], st["func_eval"])
    assert abs(fx8 - fx) <= 1.0e-10 * abs(fx)


def test_scg_variants_agree_on_random_small_problems(monkeypatch):
    """tools/scg_soak.py in small: random Lorenz-96 problems, pipelining off / default / forced
    bit-identical under both reductions, and the same branch counts across the reductions"""
    from helpers import synthetic_l96
    rng = np.random.default_rng(2026)
    for _ in range(8):
        D, M = int(rng.integers(4, 70)), int(rng.integers(3, 14))
        g = synthetic_l96(D, M, seed=int(rng.integers(1 << 30)), partial=bool(rng.integers(2)))
        iters = int(rng.integers(10, 60))
        out = {}
        for cl in ("0", "1"):
            for pipe in ("0", "1", "2"):
                monkeypatch.setenv("MFVGP_SCG_PIPELINE", pipe)
                monkeypatch.setenv("MFVGP_SCG_CLUSTER", cl)
                x, fx, st = make_free_energy(g).find_minimum(g["x0"].copy(), max_iter=iters)
                out[cl, pipe] = (x, fx, int(st["nit"]), int(st["func_eval"]))
        for cl in ("0", "1"):
            for pipe in ("1", "2"):
                a, b = out[cl, "0"], out[cl, pipe]
                assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], (D, M, cl, pipe)
        assert out["0", "0"][2:] == out["1", "0"][2:], (D, M)
