"""
CPU tests (no GPU): the C-ABI library loads and exports every symbol that
include/mfvgp.h declares; host-side logic (argument validation mirroring the
reference's constructors, interval sharding, failing loudly without a GPU).
No compute entry is called here.
"""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from meanfieldvargp_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "mfvgp.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfvgp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from meanfieldvargp_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mfvgp.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.mfvgp_version() == 100


def test_no_torch_types_in_abi():
    text = (ROOT / "include" / "mfvgp.h").read_text()
    assert "torch" not in text.lower() and "at::" not in text


def test_create_validates_arguments(lib):
    h = ctypes.c_void_p()
    # lorenz_96.py:60-63 -- D < 4 is rejected
    assert lib.mfvgp_create(ctypes.byref(h), 3, 3, 10, 3, 0, 0, -1) == -1
    assert b"Insufficient state vector dimensions" in lib.mfvgp_last_error()
    assert lib.mfvgp_create(ctypes.byref(h), 0, 2, 10, 1, 0, 0, -1) == -1      # DW needs D=1
    assert lib.mfvgp_create(ctypes.byref(h), 9, 4, 10, 4, 0, 0, -1) == -1      # unknown system
    assert lib.mfvgp_create(ctypes.byref(h), 3, 8, 10, 9, 0, 0, -1) == -1      # d > D
    assert lib.mfvgp_create(ctypes.byref(h), 3, 8, 10, 8, 0, 5, 3) == -1       # bad shard


def test_fails_loudly_without_gpu(lib):
    if lib.mfvgp_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.mfvgp_create(ctypes.byref(h), 3, 8, 10, 8, 0, 0, -1) == -2
    assert b"no CPU fallback" in lib.mfvgp_last_error()
    import meanfieldvargp_b200 as mf
    sde = mf.Lorenz96(40.0, 8.0, dim_D=8)
    sde.time_window = np.array([0.0, 1.0])
    with pytest.raises(mf.MfvgpError, match="no CUDA device"):
        mf.FreeEnergy(sde, np.ones(8), np.ones(8), np.arange(1, 5) * 0.25,
                      np.zeros((4, 8)), 2.0 * np.ones(8))


def test_product_never_imports_oracle():
    for p in (ROOT / "meanfieldvargp_b200").rglob("*.py"):
        assert "oracle" not in p.read_text(), f"{p} mentions the oracle"


def test_system_holders_mirror_reference_validation():
    import meanfieldvargp_b200 as mf
    with pytest.raises(ValueError):
        mf.Lorenz96(40.0, 8.0, dim_D=3)                      # test_lorenz_96.py:42-63
    with pytest.raises(ValueError):
        mf.DoubleWell(-1.0, 1.0)                             # test_stochastic_process.py
    with pytest.raises(RuntimeError):
        mf.Lorenz63([1.0, 1.0, 1.0], [10.0, 28.0])
    s = mf.Lorenz96(2.0, 8.0, dim_D=5)
    assert s.sigma.shape == (5,) and np.allclose(s.inverse_sigma, 0.5)
    with pytest.raises(NotImplementedError):
        s.time_window
    assert mf.systems.system_of(s) == "L96"


def test_shard_intervals_partition():
    from meanfieldvargp_b200.parallel import shard_intervals
    for L in (1, 7, 15, 4096):
        for world in (1, 2, 3, 8):
            if world > L:
                continue
            parts = [shard_intervals(L, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == L
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_reference_objects_are_accepted_by_name():
    """
    The drop-in takes the reference's own DoubleWell / OrnsteinUhlenbeck / Lorenz63 / Lorenz96
    instances (selected by class name, theta / sigma read through their accessors); anything
    else is refused loudly -- there is no silent fallback.  Live check where /root/reference
    exists (this container); the name table is checked everywhere.
    """
    import meanfieldvargp_b200 as mf
    from meanfieldvargp_b200.systems import system_of

    class Lorenz96(object):                      # a foreign class with the reference's name
        theta, sigma = np.array(8.0), 40.0 * np.ones(5)
    assert system_of(Lorenz96()) == "L96"

    class Brusselator(object):
        theta, sigma = np.array(1.0), np.ones(2)
    with pytest.raises(NotImplementedError, match="no CUDA drift plug-in"):
        system_of(Brusselator())

    from oracle.ref_harness import harness
    if not harness.available():
        pytest.skip("reference not present (GPU box)")
    cls = harness.reference_classes()
    ref = cls["Lorenz96"](40.0, 8.0, dim_D=6, r_seed=1)
    assert system_of(ref) == "L96" and np.allclose(ref.sigma, 40.0) and float(ref.theta) == 8.0
    assert system_of(cls["DoubleWell"](0.8, 1.0, r_seed=1)) == "DW"
    assert system_of(cls["Lorenz63"]([40, 40, 40], [10, 28, 2.66667], 1)) == "L63"
    assert system_of(cls["OrnsteinUhlenbeck"](1.0, [2.0, 0.5], r_seed=1)) == "OU"
