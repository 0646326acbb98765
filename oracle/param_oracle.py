"""
TEST INFRASTRUCTURE -- CPU ORACLE (numpy restatement).  Not a product path.

Gradients of the free energy with respect to the drift parameters ``theta`` and the diffusion
coefficients ``Sigma``.  The reference has the hooks for optimising
them -- the setters ``FreeEnergy.sde_theta`` / ``sde_sigma`` (free_energy.py:241-279) and the
closing note of ``E_cost`` (:745-748) -- and no implementation.  What it DOES define is the
function being differentiated: only ``E_sde`` depends on theta / Sigma (``E_kl0`` :281-336 and
``E_obs`` :567-649 do not), and (free_energy.py:423-458, 490-493)

    Esde = sum_n 1/2 sum_i (1/Sigma_i) int_n E_i(t) dt

with the integrands ``E_i`` of ``energy_functions/*.sym``.  So

    dF/dtheta_k = 1/2 sum_n sum_i (1/Sigma_i) int_n dE_i/dtheta_k dt
    dF/dSigma_j = 1/2 sum_n [ -(1/Sigma_j^2) int_n E_j dt + sum_i (1/Sigma_i) int_n dE_i/dSigma_j dt ]

and ``dE_i/dSigma_j`` is non-zero only for j = i: ``dE_i/dSigma_i = -dE_i/d(ds_i)``.  The
derivative integrands below are the closed forms of those derivatives; each integral is taken
by ``scipy.integrate.quad_vec`` with the arguments of free_energy.py:405-417, as the reference
integrates its own gradient integrands.

PARITY PINNING: ``oracle/ref_harness/make_golden_params.py`` differentiates the reference's own
decoded ``.sym`` expressions with SymPy, integrates them the same way, and also takes central
differences of the UNMODIFIED reference's ``E_cost`` through its setters; both are committed in
``tests/golden/params_*.npz`` and ``tests/test_params.py`` checks this oracle against them.
"""
import numpy as np
from scipy.integrate import quad_vec

from .mfvgp_oracle import GK21_V, GK21_X, IntervalIntegrands, _bcast

NTHETA = {"DW": 1, "OU": 2, "L63": 3, "L96": 1}


def nodal_dparams(system, m, dm, s, ds, sigma, theta):
    """dE_i/dtheta_k [D, NTH, ...] and dE_i/dSigma_i [D, ...] at the nodes (inputs [D, ...])."""
    Sig = _bcast(sigma, m)
    q = ds - Sig
    dEsig = -q / (2.0 * s)                                  # d/dSigma of q^2/(4 s)
    if system == "OU":                                      # f = theta (mu - x)
        th, mu = theta[0], theta[1]
        R = th * (mu - m[0]) - dm[0]
        dEth = np.stack((2.0 * R * (mu - m[0]) + 2.0 * th * s[0] + q[0], 2.0 * R * th))[None]
        dEsig = dEsig - th                                  # -q E[f'] with E[f'] = -theta
    elif system == "DW":                                    # f = 4 x (theta - x^2)
        th = theta[0]
        m0, s0 = m[0], s[0]
        dEf2 = 16.0 * (2.0 * th * (m0 ** 2 + s0) - 2.0 * (m0 ** 4 + 6.0 * m0 ** 2 * s0
                                                          + 3.0 * s0 ** 2))
        dEth = (dEf2 - 8.0 * m0 * dm[0] - 4.0 * q[0])[None, None]
        dEsig = dEsig + (4.0 * th - 12.0 * (m0 ** 2 + s0))  # + E[f']
    elif system == "L63":                                   # theta = (sigma, rho, beta)
        a, r, b = theta
        R0 = a * (m[1] - m[0]) - dm[0]
        R1 = m[0] * (r - m[2]) - m[1] - dm[1]
        R2 = m[0] * m[1] - b * m[2] - dm[2]
        z = np.zeros_like(m[0])
        dEth = np.stack((np.stack((2.0 * R0 * (m[1] - m[0]) + 2.0 * a * (s[0] + s[1]) + q[0], z, z)),
                         np.stack((z, 2.0 * R1 * m[0] + 2.0 * (r - m[2]) * s[0], z)),
                         np.stack((z, z, -2.0 * R2 * m[2] + 2.0 * b * s[2] + q[2]))))
        dEsig = dEsig - _bcast(np.array([a, 1.0, b]), m)
    elif system == "L96":                                   # f_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + theta
        u = np.roll(m, -1, 0) - np.roll(m, 2, 0)
        R = u * np.roll(m, 1, 0) - m + theta[0] - dm
        dEth = (2.0 * R)[:, None]
        dEsig = dEsig - 1.0
    else:
        raise ValueError(system)
    return dEth, dEsig


def param_gradients(fe, x, adaptive=True):
    """
    ``fe``: a ``FreeEnergyOracle``.  Returns (dF/dtheta [ntheta], dF/dSigma [D]) at state ``x``.
    ``adaptive=False`` takes every integral with the fixed 42-node rule.
    """
    D, M = fe.dim_D, fe.num_M
    x = np.asarray(x, dtype=float)
    mean_pts = x[:fe.num_mp].reshape(D, 3 * M + 4)
    vars_pts = np.exp(x[fe.num_mp:].reshape(D, 2 * M + 3))
    inv_sigma = 1.0 / fe.sigma
    nth = NTHETA[fe.system]
    dth = np.zeros(nth)
    dsig = np.zeros(D)
    kw = dict(limit=100, epsabs=1.0e-06, epsrel=1.0e-06)
    for n in range(fe.obs_times.size - 1):                       # free_energy.py:490
        ti, tj = fe.obs_times[n], fe.obs_times[n + 1]
        I = IntervalIntegrands(fe.system, ti, tj, mean_pts[:, 3 * n:3 * n + 4],
                               vars_pts[:, 2 * n:2 * n + 3], fe.sigma, fe.theta)

        def f_all(t):
            _, mom = I._moments(t)
            E = I._nodal(mom)[0]
            dEth, dEsig = nodal_dparams(fe.system, *mom, fe.sigma, fe.theta)
            return E, dEth, dEsig

        if adaptive:
            i_E = quad_vec(lambda t: f_all(t)[0], ti, tj, **kw)[0]
            i_th = quad_vec(lambda t: f_all(t)[1], ti, tj, **kw)[0]
            i_sg = quad_vec(lambda t: f_all(t)[2], ti, tj, **kw)[0]
        else:
            c = 0.5 * (ti + tj)
            i_E = i_th = i_sg = 0.0
            for (a, b) in ((ti, c), (c, tj)):
                cc, hh = 0.5 * (a + b), 0.5 * (b - a)
                E, dEth, dEsig = f_all(cc + hh * GK21_X)
                i_E = i_E + hh * (E @ GK21_V)
                i_th = i_th + hh * (dEth @ GK21_V)
                i_sg = i_sg + hh * (dEsig @ GK21_V)
        dth += 0.5 * np.tensordot(inv_sigma, i_th, axes=(0, 0))
        dsig += 0.5 * (inv_sigma * i_sg - inv_sigma ** 2 * i_E)
    return dth, dsig
