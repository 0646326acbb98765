"""
Drop-in for the reference's ``FreeEnergy`` (src/var_bayes/free_energy.py:19-864)
whose hot path -- ``E_sde`` (:461-565), ``E_kl0`` (:281-336), ``E_obs``
(:567-649), ``E_cost`` (:651-751) and the SCG loop behind ``find_minimum``
(:753-862) -- runs as hand-written FP64 CUDA for sm_100a behind the C ABI of
``include/mfvgp.h``.

Same constructor, same method names, argument meaning, return shapes and
exceptions as the reference.  ``n_jobs`` is accepted and ignored (the joblib
fan-out of free_energy.py:500-508 is one kernel launch here).  numpy in ->
numpy out (copies inside the call); CUDA ``torch.float64`` tensors in ->
tensors out (zero copy, caller's current stream, no synchronisation: the
reference's ``RuntimeError`` for non-finite energies is then raised by
``FreeEnergy.check_last()``, not by the call itself).

There is no CPU fallback: without the built library or a CUDA device every
compute call raises ``MfvgpError``.
"""
import ctypes
from time import perf_counter

import numpy as np

from . import _lib
from .systems import system_of


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class FreeEnergy(object):

    MAX_CPUs = 1      # kept for interface parity (free_energy.py:22); unused

    def __init__(self, sde, mu0, tau0, obs_times, obs_values, obs_noise,
                 obs_mask=None, n_jobs=1, device=0, intervals=None):
        """
        Arguments as free_energy.py:30-55.  Extra keyword arguments:
        ``device`` (CUDA ordinal) and ``intervals=(n_begin, n_end)``, the
        shard of the L = M-1 integrated intervals this object evaluates
        (multi-GPU; default all).
        """
        self.system = system_of(sde)
        self._mu0 = np.asarray(mu0, dtype=float)
        self._tau0 = np.asarray(tau0, dtype=float)
        self.theta = np.atleast_1d(np.asarray(sde.theta, dtype=float))
        self.sigma = np.atleast_1d(np.asarray(sde.sigma, dtype=float))
        self.tk = sde.time_window
        if not np.all(np.diff(self.tk) > 0.0):                   # :83-85
            raise ValueError(f" {self.__class__.__name__}:"
                             f" Time window [t0, tf] is not increasing.")
        self.obs_times = np.asarray(obs_times, dtype=float)
        self.obs_noise = np.asarray(obs_noise, dtype=float)
        self.obs_values = np.asarray(obs_values, dtype=float)
        self.dim_D = self.sigma.size
        if obs_mask is None:                                     # :98-127
            self.obs_mask = self.dim_D * [True]
        else:
            if len(obs_mask) != self.dim_D:
                raise RuntimeError(f" {self.__class__.__name__}:"
                                   f" Observation mask mismatch {len(obs_mask)} != {self.dim_D}")
            if all(isinstance(item, bool) for item in obs_mask):
                self.obs_mask = obs_mask
            else:
                raise ValueError(f" {self.__class__.__name__}:"
                                 f" Observation mask contains non-boolean values!")
            if sum(self.obs_mask) == 0:
                raise RuntimeError(f" {self.__class__.__name__}:"
                                   f" Observation mask contains only False values.")
        self.n_jobs = int(n_jobs)                                # accepted, ignored
        self.dim_d = 1 if self.obs_values.ndim in (0, 1) else self.obs_values.shape[1]
        self.num_M = len(obs_times)
        if self.dim_d > self.dim_D:                              # :153-157
            raise RuntimeError(f" {self.__class__.__name__}: System dimensions."
                               f" {self.dim_d} should be less than, or equal,"
                               f" to {self.dim_D}.")
        if self.num_M != self.obs_values.shape[0]:               # :160-164
            raise RuntimeError(f" {self.__class__.__name__}:  System dimensions."
                               f" Observation times {self.num_M} should be equal"
                               f" to observation values {self.obs_values.shape[0]}.")
        if sum(self.obs_mask) != self.dim_d:
            raise RuntimeError(f" {self.__class__.__name__}: Observation mask selects"
                               f" {sum(self.obs_mask)} dimensions, observations have"
                               f" {self.dim_d}.")
        self.num_mp = self.dim_D * (3 * self.num_M + 4)          # :168
        self.num_sp = self.dim_D * (2 * self.num_M + 3)
        self.num_L = self.num_M - 1                              # :490
        self.ikm = [3 * i for i in range(1, self.num_M + 1)]              # :171-172
        self.iks = [2 * i for i in range(1, self.num_M + 1)]
        # mesh-grids that pick the observed rows at the observation columns (:176-177):
        # mean_points[fe.ix_ikm], vars_points[fe.ix_iks] are what E_obs takes
        self.ix_ikm = np.ix_(self.obs_mask, self.ikm)
        self.ix_iks = np.ix_(self.obs_mask, self.iks)

        # ---- device side ------------------------------------------------
        self._lib = _lib.require_device()
        self.device = int(device)
        n0, n1 = (0, -1) if intervals is None else intervals
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.mfvgp_create(ctypes.byref(self._h), _lib.SYS_ID[self.system],
                                          self.dim_D, self.num_M, self.dim_d, self.device,
                                          int(n0), int(n1)), "mfvgp_create")
        self.intervals = (int(n0), self.num_L if n1 < 0 else int(n1))
        self._params_sent = None
        self._push_obs()
        self._sync_params()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.mfvgp_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- parameter upload (re-read every call: the reference's setters at
    # free_energy.py:191-279 may replace them between calls) -----------------
    def _push_obs(self):
        t = _f64(self.obs_times)
        y = _f64(self.obs_values).reshape(self.num_M, self.dim_d)
        r = _f64(np.atleast_1d(self.obs_noise))
        if r.size == 1 and self.dim_d > 1:
            r = np.full(self.dim_d, float(r[0]))
        mask = np.ascontiguousarray(np.asarray(self.obs_mask, dtype=np.uint8))
        _lib.check(self._lib.mfvgp_set_obs(self._h, _ptr(t), _ptr(y), _ptr(r), _ptr(mask)),
                   "mfvgp_set_obs")

    def _sync_params(self):
        """
        Re-reads theta, sigma, mu0, tau0 on every call (cheap byte comparison, ~2 us);
        uploads only what changed.
        """
        key = (self.sigma.tobytes(), self.theta.tobytes(), self._mu0.tobytes(),
               self._tau0.tobytes())
        if key == self._params_sent:
            return
        if key[:2] != (self._params_sent or (None, None))[:2]:
            s, t = _f64(self.sigma), _f64(self.theta)
            _lib.check(self._lib.mfvgp_set_sde(self._h, _ptr(s), _ptr(t), t.size),
                       "mfvgp_set_sde")
        mu0 = _f64(np.broadcast_to(self._mu0, (self.dim_D,)))
        tau0 = _f64(np.broadcast_to(self._tau0, (self.dim_D,)))
        _lib.check(self._lib.mfvgp_set_prior(self._h, _ptr(mu0), _ptr(tau0)),
                   "mfvgp_set_prior")
        self._params_sent = key

    # ---- accessors, free_energy.py:181-279 ---------------------------------
    @property
    def prior_mu0(self):
        return self._mu0

    @prior_mu0.setter
    def prior_mu0(self, new_value):
        self._mu0 = np.asarray(new_value, dtype=float)

    @property
    def prior_tau0(self):
        return self._tau0

    @prior_tau0.setter
    def prior_tau0(self, new_value):
        self._tau0 = np.asarray(new_value, dtype=float)

    @property
    def sde_theta(self):
        return self.theta

    @sde_theta.setter
    def sde_theta(self, new_value):
        self.theta = np.atleast_1d(np.asarray(new_value, dtype=float))

    @property
    def sde_sigma(self):
        return self.sigma

    @sde_sigma.setter
    def sde_sigma(self, new_value):
        self.sigma = np.atleast_1d(np.asarray(new_value, dtype=float))

    @property
    def kernel_launches(self):
        return int(self._lib.mfvgp_launch_count(self._h))

    def describe(self):
        """Which kernels this object runs, e.g. {'path': 'walk', 'kernel': ..., 'te': '128'}."""
        buf = ctypes.create_string_buffer(512)
        self._lib.mfvgp_describe(self._h, buf, 512)
        return dict(kv.split("=", 1) for kv in buf.value.decode().split())

    def attach_communicator(self, rank, world):
        """
        Join this object (created with ``intervals=`` = the rank's block) to the
        ``world`` ranks of a torch.distributed job: rank 0 creates the NCCL id,
        it is broadcast through torch.distributed, and the C library opens its
        own communicator (mfvgp_comm_init).  E_cost / find_minimum then become
        collective calls.
        """
        import torch
        import torch.distributed as dist
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(self._lib.mfvgp_comm_unique_id(_ptr(ident)), "mfvgp_comm_unique_id")
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=0)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        _lib.check(self._lib.mfvgp_comm_init(self._h, int(rank), int(world), _ptr(ident)),
                   "mfvgp_comm_init")
        self.rank, self.world = int(rank), int(world)

    def refine_info(self):
        """
        Bisections performed by the adaptive quadrature in the last evaluation,
        array [L, 2] (energy, dS integrand); 0 where the fixed 42-node rule was
        proven to be quad_vec's answer.  k > 0 corresponds to scipy's
        ``neval = 21 + 42 k`` (free_energy.py:405-417).
        """
        info = np.zeros((self.num_L, 2), dtype=np.int32)
        _lib.check(self._lib.mfvgp_refine_info(self._h, _ptr(info)), "mfvgp_refine_info")
        return info

    # ---- error convention ----------------------------------------------------
    def _raise_on_status(self, status, E0=None, Esde=None, Eobs=None):
        name = self.__class__.__name__
        if status & _lib.ST_SYNC:
            raise _lib.MfvgpError("internal synchronisation timeout in the Lorenz-96 kernel "
                                  "(MFVGP_ST_SYNC_TIMEOUT): the gradient is incomplete")
        if status & _lib.ST_E0:                                   # :313-316
            raise RuntimeError(f" {name}: E0 is not a finite number: {E0}")
        if status & _lib.ST_ESDE:                                 # :547-550
            raise RuntimeError(f" {name}: Esde is not a finite number: {Esde}")
        if status & _lib.ST_EOBS:                                 # :615-618
            raise RuntimeError(f" {name}: Eobs is not a finite number: {Eobs}")

    def check_last(self):
        """
        Deferred error check of the CUDA-tensor path.  ``E_cost(tensor)`` returns device
        tensors without synchronising, so it cannot raise the reference's ``RuntimeError`` for a
        non-finite E0 / Esde / Eobs (free_energy.py:313-316, 547-550, 615-618) at call time as
        the numpy path does; this reads the status word of the last tensor evaluation (one
        synchronising 40-byte copy), raises exactly what the numpy path would have raised, and
        returns the status bits otherwise.
        """
        out = getattr(self, "last_out", None)
        if out is None:
            return 0
        host = out.detach().cpu()
        status = int(host[4:5].view(dtype=host.dtype).numpy().view(np.int32)[0])
        self.last_status = status
        if status & 39:
            self._raise_on_status(status, E0=float(host[1]), Esde=float(host[2]),
                                  Eobs=float(host[3]))
        return status

    # ---- E_kl0, free_energy.py:281-336 ---------------------------------------
    def E_kl0(self, m0, s0, output_gradients=True):
        self._sync_params()
        want = 1 if output_gradients else 0
        if _is_tensor(m0):                   # tensors in -> tensors out, caller's stream
            import torch
            m0d = m0.contiguous().reshape(-1)
            s0d = torch.as_tensor(s0, device=m0d.device).contiguous().reshape(-1)
            D = self.dim_D
            if m0d.numel() != D or s0d.numel() != D:
                raise ValueError(f"E_kl0: m0 and s0 must have {D} entries")
            res = torch.empty(2 + 2 * D, dtype=torch.float64, device=m0d.device)
            p = res.data_ptr()
            stream = torch.cuda.current_stream(m0d.device).cuda_stream
            _lib.check(self._lib.mfvgp_ekl0(self._h, m0d.data_ptr(), s0d.data_ptr(), want, p,
                                            p + 16 if want else None,
                                            p + 16 + 8 * D if want else None, p + 8, stream),
                       "mfvgp_ekl0")
            self.last_out = None
            if not output_gradients:
                return res[0], None, None
            return res[0], res[2:2 + D], res[2 + D:]
        m0h = _f64(m0).reshape(self.dim_D)
        s0h = _f64(s0).reshape(self.dim_D)
        E0, status = ctypes.c_double(), ctypes.c_int32()
        dm0 = np.empty(self.dim_D) if want else None
        ds0 = np.empty(self.dim_D) if want else None
        _lib.check(self._lib.mfvgp_ekl0_host(self._h, _ptr(m0h), _ptr(s0h), want,
                                             ctypes.byref(E0), _ptr(dm0) if want else None,
                                             _ptr(ds0) if want else None,
                                             ctypes.byref(status)), "mfvgp_ekl0_host")
        self._raise_on_status(status.value, E0=2.0 * E0.value)
        if not output_gradients:
            return E0.value, None, None
        return E0.value, dm0, ds0

    # ---- E_obs, free_energy.py:567-649 ----------------------------------------
    def E_obs(self, mean_pts, vars_pts, output_gradients=True):
        self._sync_params()
        d, M = self.dim_d, self.num_M
        if _is_tensor(mean_pts):             # tensors in -> tensors out, caller's stream
            import torch
            mo = mean_pts.contiguous().reshape(d, M)
            so = torch.as_tensor(vars_pts, device=mo.device).contiguous().reshape(d, M)
            want = 1 if output_gradients else 0
            res = torch.empty(2 + 2 * d * M, dtype=torch.float64, device=mo.device)
            p = res.data_ptr()
            stream = torch.cuda.current_stream(mo.device).cuda_stream
            _lib.check(self._lib.mfvgp_eobs(self._h, mo.data_ptr(), so.data_ptr(), want, p,
                                            p + 16 if want else None,
                                            p + 16 + 8 * d * M if want else None, p + 8, stream),
                       "mfvgp_eobs")
            if not output_gradients:
                return res[0], None, None
            dm, ds = res[2:2 + d * M].view(d, M), res[2 + d * M:].view(d, M)
            if d == 1:                                            # :637-642
                return res[0], dm.squeeze(), ds.squeeze()
            return res[0], dm, ds
        mo = _f64(mean_pts).reshape(d, M)
        so = _f64(vars_pts).reshape(d, M)
        want = 1 if output_gradients else 0
        Eobs, status = ctypes.c_double(), ctypes.c_int32()
        dm = np.empty((d, M)) if want else None
        ds = np.empty((d, M)) if want else None
        _lib.check(self._lib.mfvgp_eobs_host(self._h, _ptr(mo), _ptr(so), want,
                                             ctypes.byref(Eobs), _ptr(dm) if want else None,
                                             _ptr(ds) if want else None,
                                             ctypes.byref(status)), "mfvgp_eobs_host")
        self._raise_on_status(status.value, Eobs=2.0 * Eobs.value)
        if not output_gradients:
            return Eobs.value, None, None
        if d == 1:                                                # :637-642
            return Eobs.value, dm.squeeze(), ds.squeeze()
        return Eobs.value, dm, ds

    # ---- E_sde, free_energy.py:461-565 ---------------------------------------
    def E_sde(self, mean_pts, vars_pts, output_gradients=True):
        self._sync_params()
        D, L = self.dim_D, self.num_L
        want = 1 if output_gradients else 0
        if _is_tensor(mean_pts):
            import torch
            mp = mean_pts.contiguous()
            sp = vars_pts.contiguous()
            out = torch.empty(1, dtype=torch.float64, device=mp.device)
            st = torch.empty(1, dtype=torch.int32, device=mp.device)
            dm = torch.empty((L, D, 4), dtype=torch.float64, device=mp.device) if want else None
            ds = torch.empty((L, D, 3), dtype=torch.float64, device=mp.device) if want else None
            stream = torch.cuda.current_stream(mp.device).cuda_stream
            _lib.check(self._lib.mfvgp_esde(
                self._h, mp.data_ptr(), mp.shape[1], sp.data_ptr(), sp.shape[1], want,
                out.data_ptr(), dm.data_ptr() if want else None,
                ds.data_ptr() if want else None, st.data_ptr(), stream), "mfvgp_esde")
            self.last_status = int(st.item())
            Esde = out[0]
            self._raise_on_status(self.last_status, Esde=float(Esde))
            return Esde, dm, ds
        mp = _f64(mean_pts).reshape(D, 3 * self.num_M + 4)
        sp = _f64(vars_pts).reshape(D, 2 * self.num_M + 3)
        Esde = ctypes.c_double()
        status = ctypes.c_int32()
        dm = np.empty((L, D, 4)) if want else None
        ds = np.empty((L, D, 3)) if want else None
        _lib.check(self._lib.mfvgp_esde_host(
            self._h, _ptr(mp), _ptr(sp), want, ctypes.byref(Esde),
            _ptr(dm) if want else None, _ptr(ds) if want else None,
            ctypes.byref(status)), "mfvgp_esde_host")
        self.last_status = status.value
        self._raise_on_status(status.value, Esde=Esde.value)
        if not output_gradients:
            return Esde.value, None, None
        return Esde.value, dm, ds

    # ---- E_cost, free_energy.py:651-751 --------------------------------------
    def E_cost(self, x, output_gradients=True):
        self._sync_params()
        n = self.num_mp + self.num_sp
        want = 1 if output_gradients else 0
        if _is_tensor(x):
            import torch
            xd = x if x.is_contiguous() else x.contiguous()
            if xd.numel() != n:
                raise ValueError(f"E_cost: x has {xd.numel()} entries, expected {n}")
            dev = xd.device
            # out = [F, E0, Esde, Eobs, status bits (int32 in the 5th slot)]
            out = torch.empty(5, dtype=torch.float64, device=dev)
            grad = torch.empty(n, dtype=torch.float64, device=dev) if want else None
            stream = torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None
                                                        else torch.cuda.current_device())
            p_out = out.data_ptr()
            _lib.check(self._lib.mfvgp_ecost(self._h, xd.data_ptr(), want, p_out,
                                             grad.data_ptr() if want else None,
                                             p_out + 32, stream), "mfvgp_ecost")
            self.last_out = out
            return (out[0], grad) if output_gradients else out[0]
        xh = x if (type(x) is np.ndarray and x.dtype == np.float64 and x.ndim == 1
                   and x.flags.c_contiguous) else _f64(x).ravel()
        if xh.size != n:
            raise ValueError(f"E_cost: x has {xh.size} entries, expected {n}")
        # one page-locked buffer receives [gradient | F, E0, Esde, Eobs | status] in one DMA
        buf = _lib.pinned_empty(n + 8)
        _lib.check(self._lib.mfvgp_ecost_host_packed(self._h, xh.ctypes.data, want,
                                                     buf.ctypes.data), "mfvgp_ecost_host_packed")
        out = buf[n:n + 4]
        status = int(buf[n + 4:n + 5].view(np.int32)[0])
        self.last_status = status
        self.last_terms = {"F": out[0], "E0": out[1], "Esde": out[2], "Eobs": out[3]}
        if status & 39:
            self._raise_on_status(status, E0=out[1], Esde=out[2], Eobs=out[3])
        if not output_gradients:
            return float(out[0])
        return float(out[0]), buf[:n]

    # ---- hyper-parameter gradients: the hooks at free_energy.py:241-279, :745-748 ---------------
    def grad_params(self, x):
        """
        Gradients of ``E_cost`` with respect to the drift parameters and the diffusion
        coefficients at state ``x``: ``(dF_dtheta [len(theta)], dF_dsigma [D])``.  The
        reference keeps the ``sde_theta`` / ``sde_sigma`` setters "to be used only if we
        optimize the drift / diffusion parameters" (free_energy.py:241-279) and implements no
        gradient for them (:745-748); ``E_kl0`` and ``E_obs`` do not depend on either, so these
        are the derivatives of ``E_sde`` (csrc/eparam.cuh states the formulas).  numpy in ->
        numpy out, CUDA tensor in -> tensors out.
        """
        self._sync_params()
        n = self.num_mp + self.num_sp
        nth = self.theta.size
        if _is_tensor(x):
            import torch
            xd = x.contiguous().reshape(-1)
            if xd.numel() != n:
                raise ValueError(f"grad_params: x has {xd.numel()} entries, expected {n}")
            res = torch.empty(nth + self.dim_D + 1, dtype=torch.float64, device=xd.device)
            stream = torch.cuda.current_stream(xd.device).cuda_stream
            p = res.data_ptr()
            _lib.check(self._lib.mfvgp_eparam(self._h, xd.data_ptr(), p, p + 8 * nth,
                                              p + 8 * (nth + self.dim_D), stream), "mfvgp_eparam")
            return res[:nth], res[nth:nth + self.dim_D]
        xh = _f64(x).ravel()
        if xh.size != n:
            raise ValueError(f"grad_params: x has {xh.size} entries, expected {n}")
        dth, dsg = np.empty(nth), np.empty(self.dim_D)
        status = ctypes.c_int32()
        _lib.check(self._lib.mfvgp_eparam_host(self._h, _ptr(xh), _ptr(dth), _ptr(dsg),
                                               ctypes.byref(status)), "mfvgp_eparam_host")
        self.last_status = status.value
        if status.value & _lib.ST_ESDE:
            raise RuntimeError(f" {self.__class__.__name__}: the gradient with respect to the "
                               f"drift / diffusion parameters is not finite.")
        return dth, dsg

    # ---- batched evaluation (no counterpart call in the reference; see check_gradients) ----------
    def supports_batch(self):
        """True where mfvgp_ecost_batch applies: the small-problem kernels (DW, OU, L63, Lorenz 96
        up to 37 888 cells; MFVGP_COOP=0 switches them off) on one GPU."""
        import os
        return (getattr(self, "world", 1) == 1 and os.environ.get("MFVGP_COOP", "1")[:1] != "0"
                and self.describe().get("path") in ("coop", "small"))

    def E_cost_batch(self, X, output_gradients=True, chunk=1024):
        """
        ``E_cost`` of B independent states ``X [B, len(x)]`` of this problem in two launches per
        chunk: ``(F [B], grad [B, len(x)])`` or ``F [B]``.  CUDA ``torch.float64`` tensor in ->
        tensors out (numpy in -> numpy out).  Rows on which scipy's ``quad_vec`` would bisect
        further (guard flagged, MFVGP_ST_REFINED) are re-evaluated one by one with the adaptive
        stage, so every row equals what ``E_cost`` returns for it.  What
        ``find_minimum(check_gradients=True)`` uses for its len(x) + 1 evaluations.
        """
        import torch
        self._sync_params()
        n = self.num_mp + self.num_sp
        if not self.supports_batch():
            raise NotImplementedError("E_cost_batch: small-problem path on one GPU only")
        numpy_in = not _is_tensor(X)
        Xd = torch.as_tensor(_f64(X), device="cuda") if numpy_in else X
        if Xd.dim() != 2 or Xd.shape[1] != n or Xd.dtype != torch.float64:
            raise ValueError(f"E_cost_batch: X must be float64 [B, {n}]")
        Xd = Xd.contiguous()
        B = Xd.shape[0]
        want = 1 if output_gradients else 0
        out = torch.empty((B, 8), dtype=torch.float64, device=Xd.device)
        grad = torch.empty((B, n), dtype=torch.float64, device=Xd.device) if want else None
        status = torch.empty(B, dtype=torch.int32, device=Xd.device)
        stream = torch.cuda.current_stream(Xd.device).cuda_stream
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            _lib.check(self._lib.mfvgp_ecost_batch(
                self._h, nb, Xd[b0].data_ptr(), want, out[b0].data_ptr(),
                grad[b0].data_ptr() if want else None, status[b0].data_ptr(), stream),
                "mfvgp_ecost_batch")
        st = status.cpu().numpy()
        F = out[:, 0].clone()
        for b in np.nonzero(st & _lib.ST_REFINED)[0]:          # adaptive stage, one by one
            if want:
                fb, gb = self.E_cost(Xd[int(b)])
                grad[int(b)] = gb
            else:
                fb = self.E_cost(Xd[int(b)], output_gradients=False)
            F[int(b)] = fb
            torch.cuda.synchronize()
            st[b] = int(self.last_out[4:5].view(torch.int32)[0].item()) | (st[b] & ~_lib.ST_REFINED)
        bad = np.nonzero(st & 39)[0]
        if bad.size:
            self._raise_on_status(int(st[bad[0]]), E0=float(out[bad[0], 1]),
                                  Esde=float(out[bad[0], 2]), Eobs=float(out[bad[0], 3]))
        if numpy_in:
            F = F.cpu().numpy()
            return (F, grad.cpu().numpy()) if output_gradients else F
        return (F, grad) if output_gradients else F

    def check_grad_batched(self, x, chunk=512):
        """
        ``scipy.optimize.check_grad(E_cost, grad, x)`` (free_energy.py:805, :858) -- the 2-norm of
        the analytic gradient minus forward differences with scipy's step sqrt(eps) -- with the
        len(x) perturbed states evaluated in batches on the device instead of one call each.
        """
        import torch
        x = np.asarray(x, dtype=float).ravel()
        n = x.size
        eps = float(np.sqrt(np.finfo(float).eps))
        xd = torch.as_tensor(x, device="cuda")
        _, g = self.E_cost(xd)
        self.check_last()
        fd = torch.empty(n, dtype=torch.float64, device="cuda")
        for i0 in range(0, n, chunk):
            m = min(chunk, n - i0)
            X = xd.repeat(m + 1, 1)       # last row: x itself, through the same kernels as the rest
            idx = torch.arange(m, device="cuda")
            X[idx, i0 + idx] = X[idx, i0 + idx] + eps
            h = X[idx, i0 + idx] - xd[i0:i0 + m]          # the step actually taken (scipy does the same)
            F = self.E_cost_batch(X, output_gradients=False, chunk=m + 1)
            fd[i0:i0 + m] = (F[:m] - F[m]) / h
        return float(torch.sqrt(torch.sum((g - fd) ** 2)).item())

    # ---- find_minimum, free_energy.py:753-862 --------------------------------
    def find_minimum(self, x0, max_iter=100, x_tol=1.0e-5, f_tol=1.0e-5,
                     check_gradients=False, verbose=False):
        self._sync_params()

        def _grad_check(tag, xin):
            from scipy.optimize import check_grad
            print(f"Grad-Check |{tag}| minimization ...")
            if self.supports_batch():
                err = self.check_grad_batched(np.array(xin, dtype=float))
            else:
                err = check_grad(lambda v: self.E_cost(v, output_gradients=False),
                                 lambda v: self.E_cost(v)[1], np.array(xin, dtype=float))
            print(f" > Error = {err:.3E}")
            print("------------------------------------\n")

        tensor_in = _is_tensor(x0)
        if check_gradients:
            _grad_check("BEFORE", x0.detach().cpu().numpy() if tensor_in else x0)
        max_iter = max(int(max_iter), 10)                          # :815-817
        x_tol = max(float(x_tol), 1.0e-20)
        f_tol = max(float(f_tol), 1.0e-20)
        if verbose:
            print("SCG: Optimization started ...")
        time_t0 = perf_counter()
        opt_x, opt_fx, stats = self._scg(x0, max_iter, x_tol, f_tol)
        time_tf = perf_counter()
        nit, fx_hist, dfx_hist = stats["nit"], stats["fx"], stats["dfx"]
        if nit == max_iter:                                        # scaled_cg.py:376-385
            print(f"SGC: Maximum number of iterations ({max_iter}) has been reached.")
        if verbose:
            for j in range(0, nit, 50):                           # scaled_cg.py:288-296
                print(f"It= {j:>5}: F(x)= {fx_hist[j]:.3E} -:- "
                      f"Sum(|Gradients|)= {dfx_hist[j]:.3E}")
        print(f"Elapsed time [{nit}]: {(time_tf - time_t0):.2f} seconds.\n")
        if check_gradients:
            _grad_check("AFTER", opt_x.detach().cpu().numpy() if tensor_in else opt_x)
        print("Done!")
        return opt_x, opt_fx, stats

    def _scg(self, x0, max_iter, x_tol, f_tol):
        """SCG.minimize over E_cost on the device (mfvgp_scg): (x, fx, stats)."""
        self._sync_params()
        tensor_in = _is_tensor(x0)
        fx_hist = np.zeros(max_iter)
        dfx_hist = np.zeros(max_iter)
        result = np.zeros(4)
        n = self.num_mp + self.num_sp
        if tensor_in:
            import torch
            opt_x = x0.detach().clone().contiguous().reshape(-1)
            if opt_x.numel() != n:
                raise ValueError(f"find_minimum: x0 has {opt_x.numel()} entries, expected {n}")
            stream = torch.cuda.current_stream(opt_x.device).cuda_stream
            _lib.check(self._lib.mfvgp_scg(self._h, opt_x.data_ptr(), max_iter, x_tol, f_tol,
                                           _ptr(fx_hist), _ptr(dfx_hist), _ptr(result),
                                           stream), "mfvgp_scg")
        else:
            opt_x = np.array(x0, dtype=np.float64).ravel()
            if opt_x.size != n:
                raise ValueError(f"find_minimum: x0 has {opt_x.size} entries, expected {n}")
            _lib.check(self._lib.mfvgp_scg_host(self._h, _ptr(opt_x), max_iter, x_tol, f_tol,
                                                _ptr(fx_hist), _ptr(dfx_hist), _ptr(result)),
                       "mfvgp_scg_host")
        opt_fx, nit, func_eval, status = (float(result[0]), int(result[1]),
                                          int(result[2]), int(result[3]))
        self.last_status = status
        self._raise_on_status(status, E0=opt_fx, Esde=opt_fx, Eobs=opt_fx)
        stats = {"nit": nit, "fx": fx_hist, "dfx": dfx_hist, "func_eval": func_eval}
        return opt_x, opt_fx, stats
