"""
Posterior reconstruction and the wire format of the optimisation vector -- the steps either
side of the free-energy path in the reference's notebooks.

``reconstruct`` replaces the triple Python loop of examples/example_500D.ipynb cell 29
(:537-575), which calls the two lambdified Lagrange polynomials of
``symbolics.get_local_polynomials`` (src/numerical/symbolics.py:102-149) once per
(time sample, dimension), by one CUDA kernel (csrc/posterior.cuh).  ``support_indices``,
``pack`` and ``unpack`` are the host-side bookkeeping of cells 17/19 and of
free_energy.py:678-684; ``initial_path`` / ``initial_state`` are cells 14 and 19 (the start
of the optimisation; scipy on the host, as in the notebook).  No CPU fallback: ``reconstruct``
needs the library and a GPU.
"""
import ctypes

import numpy as np

from . import _lib


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def time_knots(t0, obs_times, tf):
    """Tx = [t0, *obs_times, tf] (example_500D.ipynb:543)."""
    return np.concatenate(([float(t0)], np.asarray(obs_times, dtype=float).ravel(), [float(tf)]))


def reconstruct(mean_pts, vars_pts, t0, obs_times, tf, num=100, log_vars=False, device=0):
    """
    m(t), s(t) on the fine grid of the notebooks: every interval of
    ``Tx = [t0, *obs_times, tf]`` sampled at ``np.linspace(ti, tj, num, endpoint=False)``,
    plus ``tf``.  ``mean_pts`` [D, 3M+4], ``vars_pts`` [D, 2M+3] (the optimised support
    points; pass ``log_vars=True`` to hand over the log-variance block of ``x`` directly).

    Returns ``(tt [P], mt [D, P], st [D, P])`` with ``P = (M+1) num + 1`` -- numpy for numpy
    inputs, CUDA tensors (same stream, no copies) for CUDA ``torch.float64`` inputs.
    """
    Tx = time_knots(t0, obs_times, tf)
    M = Tx.size - 2
    num = int(num)
    if num < 1:
        raise ValueError("reconstruct: num must be >= 1")
    P = (M + 1) * num + 1
    lib = _lib.require_device()
    if _is_tensor(mean_pts):
        import torch
        mp = mean_pts if mean_pts.is_contiguous() else mean_pts.contiguous()
        sp = vars_pts if vars_pts.is_contiguous() else vars_pts.contiguous()
        D = mp.shape[0]
        if mp.shape[1] != 3 * M + 4 or sp.shape != (D, 2 * M + 3):
            raise ValueError(f"reconstruct: expected [{D}, {3 * M + 4}] mean and "
                             f"[{D}, {2 * M + 3}] variance points for M = {M}")
        Tx_d = torch.tensor(Tx, dtype=torch.float64, device=mp.device)
        tt = torch.empty(P, dtype=torch.float64, device=mp.device)
        mt = torch.empty((D, P), dtype=torch.float64, device=mp.device)
        st = torch.empty((D, P), dtype=torch.float64, device=mp.device)
        stream = torch.cuda.current_stream(mp.device).cuda_stream
        _lib.check(lib.mfvgp_reconstruct(D, M, num, Tx_d.data_ptr(), mp.data_ptr(), mp.shape[1],
                                         sp.data_ptr(), sp.shape[1], 1 if log_vars else 0,
                                         tt.data_ptr(), mt.data_ptr(), st.data_ptr(), stream),
                   "mfvgp_reconstruct")
        return tt, mt, st
    mp = np.ascontiguousarray(np.atleast_2d(np.asarray(mean_pts, dtype=np.float64)))
    sp = np.ascontiguousarray(np.atleast_2d(np.asarray(vars_pts, dtype=np.float64)))
    D = mp.shape[0]
    if mp.shape[1] != 3 * M + 4 or sp.shape != (D, 2 * M + 3):
        raise ValueError(f"reconstruct: expected [{D}, {3 * M + 4}] mean and "
                         f"[{D}, {2 * M + 3}] variance points for M = {M}")
    tt, mt, st = np.empty(P), np.empty((D, P)), np.empty((D, P))
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)                    # noqa: E731
    _lib.check(lib.mfvgp_reconstruct_host(int(device), D, M, num, ptr(Tx), ptr(mp), ptr(sp),
                                          1 if log_vars else 0, ptr(tt), ptr(mt), ptr(st)),
               "mfvgp_reconstruct_host")
    return tt, mt, st


def support_indices(obs_idk, n_t):
    """
    Time-grid indices of the mean / variance support points (example_500D.ipynb cell 17,
    :356-384): every interval of ``[0, *obs_idk, n_t - 1]`` contributes 4 (mean) and 3
    (variance) equally spaced integer positions, duplicates at the observation times removed.
    """
    time_space = [0, *[int(i) for i in obs_idk], int(n_t) - 1]
    mp_idk, sp_idk = [], []
    for n in range(len(time_space) - 1):
        mp_idk.extend(np.linspace(time_space[n], time_space[n + 1], num=4, endpoint=True, dtype=int))
        sp_idk.extend(np.linspace(time_space[n], time_space[n + 1], num=3, endpoint=True, dtype=int))
    return np.unique(mp_idk), np.unique(sp_idk)


def pack(mean_pts, vars_pts):
    """x = [means | log-variances], both row-major [D][cols] (free_energy.py:678-684;
    the notebooks' ``np.concatenate((mp_t0.flatten(), sp_t0.flatten()))`` with sp in log-space)."""
    return np.concatenate((np.asarray(mean_pts, dtype=float).ravel(),
                           np.log(np.asarray(vars_pts, dtype=float)).ravel()))


def unpack(x, dim_D, num_M):
    """Inverse of :func:`pack`: ``(mean_pts [D, 3M+4], vars_pts [D, 2M+3])`` (variances)."""
    x = np.asarray(x, dtype=float).ravel()
    nm, ns = dim_D * (3 * num_M + 4), dim_D * (2 * num_M + 3)
    if x.size != nm + ns:
        raise ValueError(f"unpack: x has {x.size} entries, expected {nm + ns}")
    return x[:nm].reshape(dim_D, 3 * num_M + 4), np.exp(x[nm:]).reshape(dim_D, 2 * num_M + 3)


def initial_path(tk, obs_idk, obs_val, h_mask=None, sample_path=None, window=51, polyorder=3,
                 tf=None, dt=None):
    """
    The notebooks' initial sample path ``x0(t)`` [D, len(tk)] (example_500D.ipynb cell 14,
    :356-400): observed dimensions get the cubic B-spline through ``[obs[0], *obs, obs[-1]]`` at
    ``[t0, *tk[obs_idk], tf]`` (``scipy.interpolate.splrep(k=3)`` / ``splev``), unobserved ones the
    Savitzky-Golay smoothing of ``sample_path[:, i]`` (``savgol_filter(.., 51, 3)``).

    :param obs_val: [M, d] noisy observations (d = number of observed dimensions)
    :param h_mask: [D] booleans (default: all observed)
    :param sample_path: [len(tk), D], needed only when some dimension is unobserved
    :param tf, dt: the notebook's window end and step.  Its spline knots are
                   ``[t0, *(obs_idk * dt), tf]`` with the NOMINAL ``tf`` -- ``tk`` itself runs one
                   step further when ``np.arange(t0, tf + dt, dt)`` rounds up (T=[0,4], dt=1e-3:
                   4 002 points, tk[-1] = 4.001) -- so pass both to reproduce the notebook's path
                   (tests/golden/initial_*.npz); default: ``tk[-1]`` and ``tk[obs_idk]``.
    """
    from scipy.interpolate import splev, splrep
    from scipy.signal import savgol_filter
    tk = np.asarray(tk, dtype=float)
    obs_val = np.asarray(obs_val, dtype=float)
    if obs_val.ndim == 1:
        obs_val = obs_val[:, None]
    D = len(h_mask) if h_mask is not None else obs_val.shape[1]
    mask = [True] * D if h_mask is None else list(h_mask)
    idk = np.asarray(obs_idk, dtype=int)
    spl_timex = np.array([tk[0], *(tk[idk] if dt is None else idk * dt),
                          tk[-1] if tf is None else tf])
    spl_value = np.array([obs_val[0], *obs_val, obs_val[-1]])
    x0, j = [], 0
    for i in range(D):
        if mask[i]:
            yy = splev(tk, splrep(spl_timex, spl_value[:, j], k=3))
            j += 1
        else:
            if sample_path is None:
                raise ValueError("initial_path: sample_path is needed for unobserved dimensions")
            yy = savgol_filter(np.asarray(sample_path)[:, i], window, polyorder)
        x0.append(yy)
    return np.array(x0)


def initial_state(x0_path, obs_idk, rng=None, var_base=4.0, var_spread=2.0):
    """
    ``x`` for ``find_minimum`` from an initial path (example_500D.ipynb cells 17 and 19,
    :402-428): means = the path at the mean support indices, variances
    ``var_base + var_spread * U[0,1)`` at the variance support indices, in log space.

    :param rng: ``numpy.random.Generator`` (the notebook draws from the unseeded global
                ``np.random.rand``; pass a generator for a reproducible start)
    :return: ``(x, mp_idk, sp_idk)``

    A CUDA ``torch.float64`` path [D, T] stays on the device: the gather and the logs run in
    ``initial_state_kernel`` (mfvgp_initial_state) and ``x`` comes back as a CUDA tensor, ready
    for ``find_minimum``.
    """
    if _is_tensor(x0_path):
        import ctypes
        import torch
        from . import _lib
        lib = _lib.require_device()
        path = x0_path.contiguous()
        D, T = path.shape
        mp_idk, sp_idk = support_indices(obs_idk, T)
        M = len(obs_idk)
        rng = np.random.default_rng() if rng is None else rng
        u = torch.as_tensor(np.ascontiguousarray(rng.random((D, len(sp_idk)))), device=path.device)
        idk = torch.as_tensor(np.asarray(mp_idk, dtype=np.int64), device=path.device)
        x = torch.empty(D * (len(mp_idk) + len(sp_idk)), dtype=torch.float64, device=path.device)
        stream = torch.cuda.current_stream(path.device).cuda_stream
        _lib.check(lib.mfvgp_initial_state(D, M, T, path.data_ptr(), idk.data_ptr(), u.data_ptr(),
                                           float(var_base), float(var_spread), x.data_ptr(),
                                           ctypes.c_void_p(stream)), "mfvgp_initial_state")
        return x, mp_idk, sp_idk
    x0_path = np.atleast_2d(np.asarray(x0_path, dtype=float))
    mp_idk, sp_idk = support_indices(obs_idk, x0_path.shape[1])
    rng = np.random.default_rng() if rng is None else rng
    mp = x0_path[:, mp_idk]
    sp = var_base * np.ones((x0_path.shape[0], len(sp_idk)))
    sp += var_spread * rng.random(sp.shape)
    return pack(mp, sp), mp_idk, sp_idk            # pack puts the variances in log space
