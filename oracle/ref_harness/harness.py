"""
TEST INFRASTRUCTURE (oracle side) -- runs the UNMODIFIED reference.

Makes ``/root/reference`` importable in this container (Python 3.12,
NumPy 2) without editing or copying any of it:

  1. ``numpy.asfarray`` shim (also exported through PYTHONPATH/sitecustomize
     for loky workers);
  2. the module-level name ``dl_load`` of the four system modules
     (``lorenz_96.py:4`` etc.) is pointed at ``symdecode.load`` so the
     CPython-3.8 ``.sym`` pickles are decoded to source instead of executed.

Everything downstream -- ``njit``, ``_construct_functions``,
``single_interval``, ``scipy.integrate.quad_vec``, joblib, ``SCG`` -- is the
reference's own code.  Only usable where /root/reference exists (this
container); the GPU box uses the committed fixtures in tests/golden/.
"""
import os
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path(os.environ.get("MFVGP_REFERENCE_ROOT", "/root/reference"))


def available() -> bool:
    return (REFERENCE_ROOT / "src" / "var_bayes" / "free_energy.py").is_file()


def install():
    """Idempotently prepare the process (and its future children)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    # children (loky workers) pick up sitecustomize.py from this directory
    pp = os.environ.get("PYTHONPATH", "")
    if str(_HERE) not in pp.split(os.pathsep):
        os.environ["PYTHONPATH"] = str(_HERE) + (os.pathsep + pp if pp else "")
    if str(_HERE) not in sys.path:
        sys.path.insert(0, str(_HERE))
    import sitecustomize  # noqa: F401  (ours: installs np.asfarray)
    import importlib
    importlib.reload(sitecustomize)
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))

    import symdecode
    from src.dynamical_systems import (double_well, lorenz_63, lorenz_96,
                                       ornstein_uhlenbeck)
    for mod in (double_well, lorenz_63, lorenz_96, ornstein_uhlenbeck):
        mod.dl_load = symdecode.load


def reference_classes():
    """Return the reference's own classes (imported from /root/reference)."""
    install()
    from src.dynamical_systems.double_well import DoubleWell
    from src.dynamical_systems.lorenz_63 import Lorenz63
    from src.dynamical_systems.lorenz_96 import Lorenz96
    from src.dynamical_systems.ornstein_uhlenbeck import OrnsteinUhlenbeck
    from src.numerical.scaled_cg import SCG
    from src.var_bayes.free_energy import FreeEnergy
    return dict(DoubleWell=DoubleWell, OrnsteinUhlenbeck=OrnsteinUhlenbeck,
                Lorenz63=Lorenz63, Lorenz96=Lorenz96, SCG=SCG,
                FreeEnergy=FreeEnergy)
