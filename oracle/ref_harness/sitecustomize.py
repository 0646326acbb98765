"""
TEST INFRASTRUCTURE (oracle side).  NumPy >= 2 removed ``numpy.asfarray``;
the reference imports it (``free_energy.py:7-9``, ``scaled_cg.py:4``).
This shim is loaded through PYTHONPATH so joblib/loky worker processes get
it too.  It changes nothing in /root/reference.
"""
import numpy as _np

if not hasattr(_np, "asfarray"):
    def _asfarray(a, dtype=float):
        return _np.asarray(a, dtype=dtype)
    _np.asfarray = _asfarray
