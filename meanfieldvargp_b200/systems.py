"""
Parameter holders for the four drift plug-ins, with the constructor signatures
of the reference's system classes:

  DoubleWell(sigma, theta, r_seed)          double_well.py:17-38
  OrnsteinUhlenbeck(sigma, theta, r_seed)   ornstein_uhlenbeck.py:15-36
  Lorenz63(sigma, theta, r_seed)            lorenz_63.py:42-90
  Lorenz96(sigma, theta, dim_D, r_seed)     lorenz_96.py:42-101

What the free-energy path reads is kept (``theta``, ``sigma``, ``time_window``;
stochastic_process.py:58-182), plus the data-generation step either side of it
(SURVEY.md 8f-3): ``make_trajectory`` / ``collect_obs`` / ``sample_path`` with the
reference's signatures, simulated on the device (``trajectory.py``) and stored as numpy
arrays like the reference's.  The energy / gradient expressions themselves live in the
CUDA kernels (csrc/drift.cuh, csrc/kernels.cuh).  A reference system object can be
passed to ``FreeEnergy`` directly: the kernel is selected from the class name.
"""
import numpy as np


class StochasticProcess(object):
    system = None

    def __init__(self, r_seed=None):
        self._rng = np.random.default_rng(r_seed)
        self.xt = None
        self.tk = None
        self._theta = None
        self._sigma = None

    # stochastic_process.py:58-84
    @property
    def theta(self):
        if self._theta is None:
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" Drift theta parameter(s) is not set yet.")
        return np.atleast_1d(self._theta)

    @theta.setter
    def theta(self, new_value):
        self._theta = np.asarray(new_value, dtype=float)

    # stochastic_process.py:86-118
    @property
    def sigma(self):
        if self._sigma is None:
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" SDE noise parameters have not set yet.")
        return np.atleast_1d(self._sigma)

    @sigma.setter
    def sigma(self, new_value):
        new_value = np.asarray(new_value, dtype=float)
        if np.any(new_value <= 0.0):
            raise ValueError(f" {self.__class__.__name__}:"
                             f" SDE noise vector {new_value} is not positive.")
        self._sigma = new_value

    @property
    def inverse_sigma(self):
        return 1.0 / self.sigma

    # stochastic_process.py:158-182
    @property
    def time_window(self):
        if self.tk is None:
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" Time window has not been created.")
        return self.tk

    @time_window.setter
    def time_window(self, new_value):
        self.tk = new_value

    @property
    def rng(self):
        return self._rng

    # stochastic_process.py:130-156
    @property
    def sample_path(self):
        if self.xt is None:
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" Sample path has not been created.")
        return self.xt

    @sample_path.setter
    def sample_path(self, new_value):
        self.xt = new_value

    def _store_path(self, tk, x):
        self.sample_path = x.cpu().numpy()
        self.time_window = tk.cpu().numpy()

    def collect_obs(self, n_obs, h_mask=None):
        """stochastic_process.py:211-297 (``trajectory.collect_obs`` on the stored path)."""
        from .trajectory import collect_obs
        if (self.tk is None) or (self.xt is None):
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" Sample path (or time window) have not been created.")
        if len(self.tk) != len(self.xt):
            raise RuntimeError(f" {self.__class__.__name__}:"
                               f" Sample path and time window do not have the same length.")
        return collect_obs(self.tk, self.xt, n_obs, h_mask, owner=self.__class__.__name__)


class DoubleWell(StochasticProcess):
    system = "DW"

    def __init__(self, sigma, theta, r_seed=None):
        super().__init__(r_seed=r_seed)
        self.sigma = sigma
        self.theta = theta

    def make_trajectory(self, t0, tf, dt=0.01):
        """double_well.py:40-91, on the device; draws from ``self.rng`` in the reference's order."""
        from .trajectory import simulate_dw
        self._store_path(*simulate_dw(self.sigma, self.theta, t0, tf, dt, rng=self.rng))


class OrnsteinUhlenbeck(StochasticProcess):
    system = "OU"

    def __init__(self, sigma, theta, r_seed=None):
        super().__init__(r_seed=r_seed)
        self.sigma = sigma
        self.theta = theta

    def make_trajectory(self, t0, tf, dt=0.01):
        """ornstein_uhlenbeck.py:38-79, on the device."""
        from .trajectory import simulate_ou
        self._store_path(*simulate_ou(self.sigma, self.theta, t0, tf, dt, rng=self.rng))


def _diag_sigma(cls_name, sigma, dim):
    sigma = np.asarray(sigma, dtype=float)
    if sigma.ndim == 0:
        sigma = sigma * np.ones(dim)
    elif sigma.ndim == 2:
        sigma = sigma.diagonal()
    elif sigma.ndim != 1:
        raise ValueError(f" {cls_name}: Wrong input dimensions: {sigma.ndim}")
    if len(sigma) != dim:
        raise ValueError(f" {cls_name}: Wrong matrix dimensions: {sigma.shape}")
    return sigma


class Lorenz63(StochasticProcess):
    system = "L63"

    def __init__(self, sigma, theta, r_seed=None):
        super().__init__(r_seed=r_seed)
        self.sigma = _diag_sigma(self.__class__.__name__, sigma, 3)
        theta = np.asarray(theta, dtype=float)
        if theta.size != 3:                       # lorenz_63.py:84-90
            raise RuntimeError(f" {self.__class__.__name__}: Drift vector"
                               f" {theta} is not [3 x 1].")
        self.theta = theta

    def make_trajectory(self, t0, tf, dt=0.01):
        """lorenz_63.py:102-181, on the device."""
        from .trajectory import simulate_l63
        self._store_path(*simulate_l63(self.sigma, self.theta, t0, tf, dt, rng=self.rng))


class Lorenz96(StochasticProcess):
    system = "L96"

    def __init__(self, sigma, theta, dim_D=40, r_seed=None):
        super().__init__(r_seed=r_seed)
        if dim_D < 4:                             # lorenz_96.py:60-63
            raise ValueError(f" {self.__class__.__name__}:"
                             f" Insufficient state vector dimensions: {dim_D}")
        self.dim_D = int(dim_D)
        self.sigma = _diag_sigma(self.__class__.__name__, sigma, self.dim_D)
        self.theta = np.asarray(theta, dtype=float)

    def make_trajectory(self, t0, tf, dt=0.01):
        """lorenz_96.py:103-185, on the device."""
        from .trajectory import simulate_l96, time_window
        noise = self.rng.standard_normal((self.dim_D, time_window(t0, tf, dt).size))     # :159-162
        self._store_path(*simulate_l96(self.sigma, float(self.theta[0]), self.dim_D, t0, tf, dt,
                                       noise=noise))


_BY_CLASS_NAME = {"DoubleWell": "DW", "OrnsteinUhlenbeck": "OU",
                  "Lorenz63": "L63", "Lorenz96": "L96"}


def system_of(sde):
    """Kernel family for an sde object (ours or the reference's own class)."""
    name = getattr(sde, "system", None) or _BY_CLASS_NAME.get(type(sde).__name__)
    if name is None:
        raise NotImplementedError(
            f"no CUDA drift plug-in for {type(sde).__name__}; supported: "
            f"{sorted(_BY_CLASS_NAME)}")
    return name
