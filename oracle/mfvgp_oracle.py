"""
TEST INFRASTRUCTURE -- CPU ORACLE (numpy restatement).  Not a product path.

Restates, for the free-energy hot path only, what the reference
(vrettasm/MeanFieldVarGP, read-only under /root/reference) computes:

  * ``FreeEnergy.E_kl0``          free_energy.py:281-336
  * ``FreeEnergy.single_interval``free_energy.py:338-459
  * ``FreeEnergy.E_sde``          free_energy.py:461-565
  * ``FreeEnergy.E_obs``          free_energy.py:567-649
  * ``FreeEnergy.E_cost``         free_energy.py:651-751
  * ``SCG.minimize``              scaled_cg.py:107-386
  * ``safe_log``                  utilities.py:42-77
  * the per-system integrands that the reference keeps as dill-pickled
    SymPy expressions (energy_functions/*.sym, gradient_functions/*.sym),
    here written in closed form (moments of the drift under N(m, s)):
    ``how_to_compute_energies.ipynb`` (cells with the Esde definition) and
    SURVEY.md section 8(a)-10 / Appendix A.

Third-party arithmetic on the path that is NOT under /root/reference:
``scipy.integrate.quad_vec`` (requirements.txt lists ``scipy`` unpinned;
1.18.1 is installed here).  The adaptive path below calls that very routine
with the same arguments as ``free_energy.py:405-417``; the fixed path
restates its first two panels (``_quad_vec.py::_quadrature_gk21``, nodes and
weights copied from the published 21-point Kronrod table).

PARITY PINNING: the reference ships no golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the
reference itself, run here through ``oracle/ref_harness`` and committed
as ``tests/golden/*.npz`` (generator: ``oracle/ref_harness/make_golden.py``).
``tests/test_oracle_vs_reference.py`` asserts the agreement.

Differences from the reference that are deliberate: the >99%-zero dense
``(D,4D)``/``(D,3D)`` integrand matrices of ``lorenz_96.py:308,348`` are kept
in neighbour-sparse form (zeros contribute nothing to sums or norms).
"""
import math

import numpy as np
from scipy.integrate import quad_vec

# --------------------------------------------------------------------------
# Gauss-Kronrod 21 tables (scipy/integrate/_quad_vec.py::_quadrature_gk21).
# x is listed descending, exactly as scipy evaluates it.
# --------------------------------------------------------------------------
_XH = (0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
       0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
       0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
       0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
       0.294392862701460198131126603103866, 0.148874338981631210884826001129720)
GK21_X = np.array(_XH + (0.0,) + tuple(-v for v in reversed(_XH)))
_VH = (0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
       0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
       0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
       0.123491976262065851077958109831074, 0.134709217311473325928054001771707,
       0.142775938577060080797094273138717, 0.147739104901338491374841515972068)
GK21_V = np.array(_VH + (0.149445554002916905664936468389821,) + tuple(reversed(_VH)))
_WH = (0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
       0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
       0.295524224714752870173892994651338)
GK21_W = np.array(_WH + tuple(reversed(_WH)))          # embedded Gauss-10

SYSTEMS = ("DW", "OU", "L63", "L96")
LOG2PI = 1.8378770664093453                            # free_energy.py:607


def safe_log(x):
    """utilities.py:73-76 -- clamp to [1e-300, 1e300], then log."""
    return np.log(np.maximum(np.minimum(np.asarray(x, dtype=float), 1.0e300),
                             1.0e-300))


# --------------------------------------------------------------------------
# Lagrange bases in ABSOLUTE time, as the reference's expressions evaluate
# them (symbolics.py:64-90): knots h_k, value and time derivative.
# --------------------------------------------------------------------------
def lagrange_basis(t, knots):
    """Return (B[k, ...], dB[k, ...]) for every knot k at times ``t``."""
    t = np.asarray(t, dtype=float)
    K = len(knots)
    diff = [t - knots[l] for l in range(K)]                  # (t - h_l)
    B = np.empty((K,) + t.shape)
    dB = np.empty((K,) + t.shape)
    for k in range(K):
        others = [l for l in range(K) if l != k]
        den = 1.0
        for l in others:
            den *= (knots[k] - knots[l])
        val = diff[others[0]]
        for l in others[1:]:
            val = val * diff[l]
        der = 0.0
        for j in others:                                    # product rule
            rest = [l for l in others if l != j]
            term = diff[rest[0]] if rest else np.ones_like(t)
            for l in rest[1:]:
                term = term * diff[l]
            der = der + term
        B[k], dB[k] = val / den, der / den
    return B, dB


def interval_knots(ti, tj):
    """free_energy.py:386-398 -- mean knots (h) and variance knots (c)."""
    delta_t = abs(tj - ti)
    h = float(delta_t / 3.0)
    c = float(delta_t / 2.0)
    return (ti, ti + h, ti + 2 * h, tj), (ti, ti + c, tj)


# --------------------------------------------------------------------------
# Per-system nodal energies E_i and their partial derivatives with respect to
# the interpolated moments.  Inputs m, dm, s, ds: arrays [D, ...].
# Common template (how_to_compute_energies.ipynb; notes/PRE_Note_01.jpg):
#   E_i = (E[f_i]-dm_i)^2 + Var[f_i] + q_i^2/(4 s_i) - q_i E[d f_i/d x_i],
#   q_i = ds_i - Sigma_i.
# Returns E [D,...], Em [D(i), D(j), ...] = dE_i/dm_j, Edm [D,...] = dE_i/d(dm_i),
# Es [D(i), D(j), ...] = dE_i/ds_j, Eds [D,...] = dE_i/d(ds_i).
# --------------------------------------------------------------------------
def _bcast(v, ref):
    return np.asarray(v, dtype=float).reshape((-1,) + (1,) * (ref.ndim - 1))


def nodal_dense(system, m, dm, s, ds, sigma, theta):
    D = m.shape[0]
    Sig = _bcast(sigma, m)
    q = ds - Sig
    rat = q * q / (4.0 * s)                 # (ds - Sigma)^2 / (4 s)
    rat_s = -q * q / (4.0 * s * s)          # d rat / d s
    rat_ds = q / (2.0 * s)                  # d rat / d ds
    Em = np.zeros((D, D) + m.shape[1:])
    Es = np.zeros((D, D) + m.shape[1:])
    if system == "OU":                      # f = theta (mu - x)
        th, mu = theta[0], theta[1]
        R = th * (mu - m[0]) - dm[0]
        E = (R * R + th * th * s[0] + rat[0] + th * q[0])[None]
        Em[0, 0] = -2.0 * th * R
        Edm = (-2.0 * R)[None]
        Es[0, 0] = th * th + rat_s[0]
        Eds = (rat_ds[0] + th)[None]
    elif system == "DW":                    # f = 4 x (theta - x^2)
        th = theta[0]
        m0, s0, d0, q0 = m[0], s[0], dm[0], q[0]
        Ef = 4.0 * (th * m0 - m0 ** 3 - 3.0 * m0 * s0)
        Ef2 = 16.0 * (th * th * (m0 ** 2 + s0)
                      - 2.0 * th * (m0 ** 4 + 6.0 * m0 ** 2 * s0 + 3.0 * s0 ** 2)
                      + m0 ** 6 + 15.0 * m0 ** 4 * s0 + 45.0 * m0 ** 2 * s0 ** 2
                      + 15.0 * s0 ** 3)
        Edf = 4.0 * th - 12.0 * (m0 ** 2 + s0)
        E = (Ef2 - 2.0 * Ef * d0 + d0 * d0 + rat[0] - q0 * Edf)[None]
        dEf_dm = 4.0 * (th - 3.0 * m0 ** 2 - 3.0 * s0)
        dEf_ds = -12.0 * m0
        dEf2_dm = 16.0 * (2.0 * th * th * m0 - 2.0 * th * (4.0 * m0 ** 3 + 12.0 * m0 * s0)
                          + 6.0 * m0 ** 5 + 60.0 * m0 ** 3 * s0 + 90.0 * m0 * s0 ** 2)
        dEf2_ds = 16.0 * (th * th - 2.0 * th * (6.0 * m0 ** 2 + 6.0 * s0)
                          + 15.0 * m0 ** 4 + 90.0 * m0 ** 2 * s0 + 45.0 * s0 ** 2)
        Em[0, 0] = dEf2_dm - 2.0 * d0 * dEf_dm + 24.0 * q0 * m0
        Edm = (-2.0 * Ef + 2.0 * d0)[None]
        Es[0, 0] = dEf2_ds - 2.0 * d0 * dEf_ds + rat_s[0] + 12.0 * q0
        Eds = (rat_ds[0] - Edf)[None]
    elif system == "L63":                   # theta = (sigma, rho, beta)
        a, r, b = theta[0], theta[1], theta[2]
        R0 = a * (m[1] - m[0]) - dm[0]
        R1 = m[0] * (r - m[2]) - m[1] - dm[1]
        R2 = m[0] * m[1] - b * m[2] - dm[2]
        V0 = a * a * (s[0] + s[1])
        V1 = m[0] ** 2 * s[2] + (r - m[2]) ** 2 * s[0] + s[0] * s[2] + s[1]
        V2 = m[0] ** 2 * s[1] + m[1] ** 2 * s[0] + s[0] * s[1] + b * b * s[2]
        E = np.stack((R0 * R0 + V0 + rat[0] + a * q[0],
                      R1 * R1 + V1 + rat[1] + q[1],
                      R2 * R2 + V2 + rat[2] + b * q[2]))
        Em[0, 0] = -2.0 * a * R0
        Em[0, 1] = 2.0 * a * R0
        Em[1, 0] = 2.0 * R1 * (r - m[2]) + 2.0 * m[0] * s[2]
        Em[1, 1] = -2.0 * R1
        Em[1, 2] = -2.0 * R1 * m[0] - 2.0 * (r - m[2]) * s[0]
        Em[2, 0] = 2.0 * R2 * m[1] + 2.0 * m[0] * s[1]
        Em[2, 1] = 2.0 * R2 * m[0] + 2.0 * m[1] * s[0]
        Em[2, 2] = -2.0 * b * R2
        Edm = np.stack((-2.0 * R0, -2.0 * R1, -2.0 * R2))
        Es[0, 0] = a * a + rat_s[0]
        Es[0, 1] = a * a + 0.0 * s[0]
        Es[1, 0] = (r - m[2]) ** 2 + s[2]
        Es[1, 1] = 1.0 + rat_s[1]
        Es[1, 2] = m[0] ** 2 + s[0]
        Es[2, 0] = m[1] ** 2 + s[1]
        Es[2, 1] = m[0] ** 2 + s[0]
        Es[2, 2] = b * b + rat_s[2]
        Eds = np.stack((rat_ds[0] + a, rat_ds[1] + 1.0, rat_ds[2] + b))
    else:
        raise ValueError(f"nodal_dense: unsupported system {system}")
    return E, Em, Edm, Es, Eds


def l96_neighbours(D):
    """lorenz_96.py:232-235 -- order (i, i+1, i-1, i-2) mod D; shape [4, D]."""
    i = np.arange(D)
    return np.stack((i, np.mod(i + 1, D), np.mod(i - 1, D), np.mod(i - 2, D)))


def nodal_l96(m, dm, s, ds, sigma, theta):
    """
    L96: f_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + theta (lorenz_96.py:9-30).
    Returns E [D,...] and neighbour-sparse partials indexed by the slot order
    of ``l96_neighbours``: Em [D, 4, ...], Edm [D, ...], Es [D, 4, ...], Eds.
    """
    th = theta[0]
    Sig = _bcast(sigma, m)
    a, b, c = np.roll(m, -1, 0), np.roll(m, 1, 0), np.roll(m, 2, 0)
    sa, sb, sc = np.roll(s, -1, 0), np.roll(s, 1, 0), np.roll(s, 2, 0)
    u = a - c
    su = sa + sc
    R = u * b - m + th - dm
    q = ds - Sig
    E = R * R + s + sb * u * u + b * b * su + su * sb + q * q / (4.0 * s) + q
    dEa = 2.0 * (R * b + sb * u)
    Em = np.stack((-2.0 * R, dEa, 2.0 * (R * u + b * su), -dEa), axis=1)
    Edm = -2.0 * R
    dEsa = b * b + sb
    Es = np.stack((1.0 - q * q / (4.0 * s * s), dEsa, u * u + su, dEsa), axis=1)
    Eds = q / (2.0 * s) + 1.0
    return E, Em, Edm, Es, Eds


# --------------------------------------------------------------------------
# Integrands of one interval, in the reference's layout minus the zeros.
# --------------------------------------------------------------------------
class IntervalIntegrands(object):
    """
    Callbacks handed to the quadrature for interval [ti, tj]
    (free_energy.py:397-417; stochastic_process.py:299-448).
    ``mp`` = mean points [D,4], ``sp`` = variance points [D,3].
    Sparse layouts: dense systems -> dM [D(i), D(j), 4], dS [D(i), D(j), 3];
    L96 -> dM [D(i), 4(slot), 4], dS [D(i), 4(slot), 3].
    """

    def __init__(self, system, ti, tj, mp, sp, sigma, theta):
        self.system = system
        self.hk, self.ck = interval_knots(ti, tj)
        self.mp = np.asarray(mp, dtype=float)
        self.sp = np.asarray(sp, dtype=float)
        self.sigma = np.atleast_1d(np.asarray(sigma, dtype=float))
        self.theta = np.atleast_1d(np.asarray(theta, dtype=float))
        self.D = self.mp.shape[0]

    def _moments(self, t):
        B3, dB3 = lagrange_basis(t, self.hk)
        B2, dB2 = lagrange_basis(t, self.ck)
        m = np.tensordot(self.mp, B3, axes=(1, 0))
        dm = np.tensordot(self.mp, dB3, axes=(1, 0))
        s = np.tensordot(self.sp, B2, axes=(1, 0))
        ds = np.tensordot(self.sp, dB2, axes=(1, 0))
        return (B3, dB3, B2, dB2), (m, dm, s, ds)

    def _nodal(self, mom):
        if self.system == "L96":
            return nodal_l96(*mom, self.sigma, self.theta)
        return nodal_dense(self.system, *mom, self.sigma, self.theta)

    def energy(self, t):
        _, mom = self._moments(t)
        return self._nodal(mom)[0]

    def all_three(self, t):
        """E [D,...], dM [D,J,4,...], dS [D,J,3,...] at times t."""
        (B3, dB3, B2, dB2), mom = self._moments(t)
        E, Em, Edm, Es, Eds = self._nodal(mom)
        dM = Em[:, :, None] * B3[None, None]
        dS = Es[:, :, None] * B2[None, None]
        own = np.arange(self.D)
        if self.system == "L96":            # slot 0 is the own dimension
            dM[:, 0] += Edm[:, None] * dB3[None]
            dS[:, 0] += Eds[:, None] * dB2[None]
        else:
            dM[own, own] += Edm[:, None] * dB3[None]
            dS[own, own] += Eds[:, None] * dB2[None]
        return E, dM, dS

    def grad_mean(self, t):
        return self.all_three(t)[1]

    def grad_variance(self, t):
        return self.all_three(t)[2]

    def contract(self, inv_sigma, i_E, i_dM, i_dS):
        """free_energy.py:423-458 -- sigma^-1 contraction and the 1/2 factors."""
        Esde = 0.5 * float(np.dot(inv_sigma, i_E))
        if i_dM is None:
            return Esde, None, None
        if self.system == "L96":
            nb = l96_neighbours(self.D)                       # [4, D]
            dm = np.zeros((self.D, 4))
            dsv = np.zeros((self.D, 3))
            for slot in range(4):
                np.add.at(dm, nb[slot], inv_sigma[:, None] * i_dM[:, slot])
                np.add.at(dsv, nb[slot], inv_sigma[:, None] * i_dS[:, slot])
        else:
            dm = np.einsum("i,ijk->jk", inv_sigma, i_dM)
            dsv = np.einsum("i,ijk->jk", inv_sigma, i_dS)
        return Esde, 0.5 * dm, 0.5 * dsv


def gk21_two_panels(f, ti, tj):
    """
    What quad_vec returns when it converges after its mandatory first
    bisection: GK21 on [ti,c] + GK21 on [c,tj] (``_quad_vec.py``
    ``_quadrature_gk`` + ``_subdivide_interval``); f is evaluated on all
    42 nodes at once (trailing axis).
    """
    c = 0.5 * (ti + tj)
    total = None
    for (a, b) in ((ti, c), (c, tj)):
        cc, hh = 0.5 * (a + b), 0.5 * (b - a)
        vals = f(cc + hh * GK21_X)                             # [..., 21]
        part = hh * np.tensordot(vals, GK21_V, axes=(-1, 0))
        total = part if total is None else total + part
    return total


def single_interval(system, ti, tj, mp, sp, sigma, theta, inv_sigma,
                    output_gradients=True, adaptive=True, info=None):
    """
    free_energy.py:338-459.  ``adaptive=True`` calls scipy's quad_vec exactly
    as the reference does; ``adaptive=False`` is the fixed 42-node rule.
    ``info`` (list) receives quad_vec's neval per integrand.
    """
    I = IntervalIntegrands(system, ti, tj, mp, sp, sigma, theta)
    kw = dict(limit=100, epsabs=1.0e-06, epsrel=1.0e-06, full_output=True)
    if adaptive:
        r = quad_vec(I.energy, ti, tj, **kw)
        i_E, nev = r[0], [r[2].neval]
        i_dM = i_dS = None
        if output_gradients:
            r = quad_vec(I.grad_mean, ti, tj, **kw)
            i_dM = r[0]
            nev.append(r[2].neval)
            r = quad_vec(I.grad_variance, ti, tj, **kw)
            i_dS = r[0]
            nev.append(r[2].neval)
        if info is not None:
            info.append(nev)
    else:
        if output_gradients:
            c = 0.5 * (ti + tj)
            i_E = i_dM = i_dS = 0.0
            for (a, b) in ((ti, c), (c, tj)):
                cc, hh = 0.5 * (a + b), 0.5 * (b - a)
                E, dM, dS = I.all_three(cc + hh * GK21_X)
                i_E = i_E + hh * (E @ GK21_V)
                i_dM = i_dM + hh * (dM @ GK21_V)
                i_dS = i_dS + hh * (dS @ GK21_V)
        else:
            i_E = gk21_two_panels(I.energy, ti, tj)
            i_dM = i_dS = None
    return I.contract(np.atleast_1d(inv_sigma), i_E, i_dM, i_dS)


# --------------------------------------------------------------------------
# The objective: E_kl0 + E_sde + E_obs and its gradient.
# --------------------------------------------------------------------------
class FreeEnergyOracle(object):
    """
    Numpy restatement of ``FreeEnergy`` (free_energy.py:19-864) for a system
    named by string.  Constructor arguments mirror free_energy.py:30-32 with
    ``sde`` replaced by (system, sigma, theta).
    """

    def __init__(self, system, sigma, theta, mu0, tau0, obs_times, obs_values,
                 obs_noise, obs_mask=None, adaptive=True, n_jobs=1):
        if system not in SYSTEMS:
            raise ValueError(f"unknown system {system}")
        self.system = system
        self.sigma = np.atleast_1d(np.asarray(sigma, dtype=float))
        self.theta = np.atleast_1d(np.asarray(theta, dtype=float))
        self.mu0 = np.asarray(mu0, dtype=float)
        self.tau0 = np.asarray(tau0, dtype=float)
        self.obs_times = np.asarray(obs_times, dtype=float)
        self.obs_noise = np.asarray(obs_noise, dtype=float)
        self.obs_values = np.asarray(obs_values, dtype=float)
        self.dim_D = self.sigma.size
        self.obs_mask = (self.dim_D * [True]) if obs_mask is None else list(obs_mask)
        self.dim_d = 1 if self.obs_values.ndim in (0, 1) else self.obs_values.shape[1]
        self.num_M = len(self.obs_times)
        self.num_mp = self.dim_D * (3 * self.num_M + 4)          # :168
        ikm = [3 * i for i in range(1, self.num_M + 1)]          # :171
        iks = [2 * i for i in range(1, self.num_M + 1)]          # :172
        self.ix_ikm = np.ix_(self.obs_mask, ikm)
        self.ix_iks = np.ix_(self.obs_mask, iks)
        self.adaptive = adaptive
        self.n_jobs = int(n_jobs)        # free_energy.py:130-133 (joblib fan-out)
        self.neval = []

    # free_energy.py:281-336
    def E_kl0(self, m0, s0):
        z0 = m0 - self.mu0
        E0 = safe_log((self.tau0 / s0).prod()) + ((z0 ** 2 + s0 - self.tau0) / self.tau0).sum()
        if not np.isfinite(E0):
            raise RuntimeError(f" FreeEnergy: E0 is not a finite number: {E0}")
        one_ = np.ones(1)
        return (0.5 * E0, np.atleast_1d(z0 / self.tau0),
                np.atleast_1d(0.5 * (one_ / self.tau0 - one_ / s0)))

    # free_energy.py:461-565
    def E_sde(self, mean_pts, vars_pts, output_gradients=True):
        D = self.dim_D
        L = self.obs_times.size - 1
        inv_sigma = np.atleast_1d(1.0 / self.sigma)
        Esde = 0.0
        dm = np.empty((L, D, 4)) if output_gradients else None
        dsv = np.empty((L, D, 3)) if output_gradients else None
        self.neval = []
        if self.n_jobs > 1:
            # free_energy.py:497-508 -- one joblib task per interval
            from joblib import Parallel, delayed
            results = Parallel(n_jobs=self.n_jobs, backend="loky")(
                delayed(single_interval)(
                    self.system, self.obs_times[n], self.obs_times[n + 1],
                    mean_pts[:, 3 * n:3 * n + 4], vars_pts[:, 2 * n:2 * n + 3],
                    self.sigma, self.theta, inv_sigma, output_gradients,
                    adaptive=self.adaptive) for n in range(L))
        else:
            results = (single_interval(
                self.system, self.obs_times[n], self.obs_times[n + 1],
                mean_pts[:, 3 * n:3 * n + 4], vars_pts[:, 2 * n:2 * n + 3],
                self.sigma, self.theta, inv_sigma, output_gradients,
                adaptive=self.adaptive, info=self.neval) for n in range(L))
        for n, (e, gm, gs) in enumerate(results):
            Esde += e
            if output_gradients:
                dm[n], dsv[n] = gm, gs
        if not np.isfinite(Esde):
            raise RuntimeError(f" FreeEnergy: Esde is not a finite number: {Esde}")
        return Esde, dm, dsv

    # free_energy.py:567-649
    def E_obs(self, mean_pts, vars_pts):
        obs_noise = np.atleast_1d(self.obs_noise)
        Ri = 1.0 / obs_noise
        Qi = 1.0 / np.sqrt(obs_noise)
        Y_minus_Hm = self.obs_values.T - mean_pts
        Z = np.multiply(Qi[:, np.newaxis], Y_minus_Hm)
        Eobs = np.sum((Z * Z).sum(axis=0) + Ri.dot(vars_pts))
        Eobs += self.num_M * (self.dim_d * LOG2PI + safe_log(obs_noise.prod()))
        if not np.isfinite(Eobs):
            raise RuntimeError(f" FreeEnergy: Eobs is not a finite number: {Eobs}")
        dEobs_dm = -np.multiply(Ri[:, np.newaxis], Y_minus_Hm)
        dEobs_ds = np.full((self.dim_d, self.num_M), 0.5 * np.atleast_2d(Ri).T, dtype=float)
        return 0.5 * Eobs, dEobs_dm, dEobs_ds

    # free_energy.py:651-751
    def E_cost(self, x, output_gradients=True):
        D, M = self.dim_D, self.num_M
        x = np.asarray(x, dtype=float)
        mean_points = x[0:self.num_mp].reshape((D, 3 * M + 4))
        vars_points = np.exp(x[self.num_mp:].reshape((D, 2 * M + 3)))
        E0, dE0_dm0, dE0_ds0 = self.E_kl0(mean_points[:, 0], vars_points[:, 0])
        Esde, dEsde_dm, dEsde_ds = self.E_sde(mean_points, vars_points, output_gradients)
        Eobs, dEobs_dm, dEobs_ds = self.E_obs(mean_points[self.ix_ikm],
                                              vars_points[self.ix_iks])
        Ecost = E0 + Esde + Eobs
        if not output_gradients:
            return Ecost
        Ecost_dm = np.zeros((D, 3 * M + 4))
        Ecost_ds = np.zeros((D, 2 * M + 3))
        Ecost_dm[:, 0:4] = dEsde_dm[0]
        Ecost_ds[:, 0:3] = dEsde_ds[0]
        for n in range(1, self.obs_times.size - 1):              # :719
            Ecost_dm[:, 3 * n] += dEsde_dm[n][:, 0]
            Ecost_ds[:, 2 * n] += dEsde_ds[n][:, 0]
            Ecost_dm[:, 3 * n + 1:3 * n + 4] = dEsde_dm[n][:, 1:]
            Ecost_ds[:, 2 * n + 1:2 * n + 3] = dEsde_ds[n][:, 1:]
        Ecost_dm[:, 0] += dE0_dm0
        Ecost_ds[:, 0] += dE0_ds0
        Ecost_dm[self.ix_ikm] += dEobs_dm
        Ecost_ds[self.ix_iks] += dEobs_ds
        np.multiply(Ecost_ds, vars_points, out=Ecost_ds)         # :743
        return Ecost, np.concatenate((Ecost_dm.ravel(), Ecost_ds.ravel()), axis=0)

    # free_energy.py:753-862 (without the prints)
    def find_minimum(self, x0, max_iter=100, x_tol=1.0e-5, f_tol=1.0e-5, events=None):
        max_iter = max(int(max_iter), 10)
        x_tol = max(float(x_tol), 1.0e-20)
        f_tol = max(float(f_tol), 1.0e-20)
        x, fx, stats = scg_minimize(self.E_cost, x0, max_iter, x_tol, f_tol, events=events)
        return x, fx, stats


# --------------------------------------------------------------------------
# Scaled conjugate gradient, scaled_cg.py:107-386 (line numbers inline).
# --------------------------------------------------------------------------
def scg_minimize(func, x0, max_it=500, x_tol=1.0e-6, f_tol=1.0e-8, events=None):
    """
    ``events`` (dict, optional) receives how often each rare branch was taken: "reject"
    (:254-279), "delta<=0" (:233-236), "mu>=0" (:194-197), "restart" (:361-364).  The reference
    keeps no such counters; goldens carry them from this restatement once its trajectory has
    been checked against the reference's (oracle/ref_harness/make_golden.py).
    """
    ev = events if events is not None else {}
    for k in ("reject", "delta<=0", "mu>=0", "restart"):
        ev[k] = 0
    stats = {"nit": max_it, "fx": np.zeros(max_it), "dfx": np.zeros(max_it),
             "func_eval": 0}
    x = np.asarray(x0, dtype=float).flatten()
    dim_x = x.size
    sigma0 = 1.0e-3
    f_now, grad_new = func(x)                                   # :145
    grad_new = np.array(grad_new, dtype=float)
    grad_old = grad_new.copy()
    stats["func_eval"] += 1
    f_old = f_now
    d = -grad_new
    success = True
    count_success = 0
    beta, beta_min, beta_max = 1.0, 1.0e-15, 1.0e100
    kappa = theta = mu = 0.0
    eps_float = np.finfo(float).eps
    for j in range(max_it):                                     # :186
        if success:
            mu = d.dot(grad_new)
            if mu >= 0.0:
                ev["mu>=0"] += 1
                d = -grad_new
                mu = d.dot(grad_new)
            kappa = d.dot(d)
            if kappa < eps_float:                               # :203
                stats["nit"] = j + 1
                return x, f_now, stats
            sigma = sigma0 / np.sqrt(kappa)
            _, g_plus = func(x + sigma * d)                     # :222
            stats["func_eval"] += 1
            theta = d.dot(g_plus - grad_new) / sigma            # :228
        delta = theta + beta * kappa
        if delta <= 0.0:
            ev["delta<=0"] += 1
            delta = beta * kappa
            beta = beta - theta / kappa
        alpha = -(mu / delta)
        x_new = x + alpha * d
        f_new, g_now = func(x_new)                              # :247
        g_now = np.array(g_now, dtype=float)
        stats["func_eval"] += 1
        Delta = 2.0 * (f_new - f_old) / (alpha * mu)
        if Delta >= 0.0:
            success = True
            count_success += 1
            f_now = f_new
            x = x_new.copy()
        else:
            ev["reject"] += 1
            success = False
            f_now = f_old
            g_now = grad_old.copy()
        total_grad = np.abs(g_now).sum()
        stats["fx"][j] = f_now
        stats["dfx"][j] = total_grad
        if success:
            if np.abs(alpha * d).max() <= x_tol and abs(f_new - f_old) <= f_tol:
                stats["nit"] = j + 1                            # :306-318
                return x, f_new, stats
            f_old = f_new
            grad_old = grad_new.copy()
            f_now, grad_new = func(x)                           # :327
            grad_new = np.array(grad_new, dtype=float)
            stats["func_eval"] += 1
            if np.isclose(grad_new.dot(grad_new), 0.0):         # :333
                stats["nit"] = j + 1
                return x, f_now, stats
        if Delta < 0.25:
            beta = min(4.0 * beta, beta_max)
        if Delta > 0.75:
            beta = max(0.5 * beta, beta_min)
        if count_success == dim_x:
            ev["restart"] += 1
            d = -grad_new
            count_success = 0
        elif success:
            gamma = max(grad_new.dot(grad_old - grad_new) / mu, 0.0)
            d = gamma * d - grad_new
    return x, f_old, stats                                      # :376-385
