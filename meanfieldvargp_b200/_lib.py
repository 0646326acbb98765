"""
ctypes binding of the C ABI declared in ``include/mfvgp.h`` (built in-tree as
``meanfieldvargp_b200/lib/libmfvgp.so`` by ``make`` / ``__graft_entry__.build``).

There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int32, c_int64,
                    c_uint8, c_void_p)
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libmfvgp.so"

SYS_ID = {"DW": 0, "OU": 1, "L63": 2, "L96": 3}
ST_ESDE, ST_E0, ST_EOBS, ST_REFINED, ST_LIMIT, ST_SYNC = 1, 2, 4, 8, 16, 32

_P = c_void_p     # every data pointer crosses the ABI as a plain address

# name -> (restype, argtypes); kept in one table so tests can check that the
# library exports every symbol the header declares
SIGNATURES = {
    "mfvgp_version": (c_int, []),
    "mfvgp_last_error": (c_char_p, []),
    "mfvgp_device_count": (c_int, []),
    "mfvgp_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mfvgp_destroy": (c_int, [c_void_p]),
    "mfvgp_set_sde": (c_int, [c_void_p, _P, _P, c_int]),
    "mfvgp_set_prior": (c_int, [c_void_p, _P, _P]),
    "mfvgp_set_obs": (c_int, [c_void_p, _P, _P, _P, _P]),
    "mfvgp_esde": (c_int, [c_void_p, _P, c_int64, _P, c_int64, c_int, _P, _P, _P, _P, c_void_p]),
    "mfvgp_ecost": (c_int, [c_void_p, _P, c_int, _P, _P, _P, c_void_p]),
    "mfvgp_esde_host": (c_int, [c_void_p, _P, _P, c_int, _P, _P, _P, _P]),
    "mfvgp_ecost_batch": (c_int, [c_void_p, c_int, _P, c_int, _P, _P, _P, c_void_p]),
    "mfvgp_ecost_host": (c_int, [c_void_p, _P, c_int, _P, _P, _P]),
    "mfvgp_ecost_host_packed": (c_int, [c_void_p, _P, c_int, _P]),
    "mfvgp_host_alloc": (c_void_p, [c_int64]),
    "mfvgp_host_free": (None, [c_void_p]),
    "mfvgp_ekl0": (c_int, [c_void_p, _P, _P, c_int, _P, _P, _P, _P, c_void_p]),
    "mfvgp_eobs": (c_int, [c_void_p, _P, _P, c_int, _P, _P, _P, _P, c_void_p]),
    "mfvgp_ekl0_host": (c_int, [c_void_p, _P, _P, c_int, POINTER(c_double), _P, _P, POINTER(c_int32)]),
    "mfvgp_eobs_host": (c_int, [c_void_p, _P, _P, c_int, POINTER(c_double), _P, _P, POINTER(c_int32)]),
    "mfvgp_eparam": (c_int, [c_void_p, _P, _P, _P, _P, c_void_p]),
    "mfvgp_eparam_host": (c_int, [c_void_p, _P, _P, _P, POINTER(c_int32)]),
    "mfvgp_scg": (c_int, [c_void_p, _P, c_int, c_double, c_double, _P, _P, _P, c_void_p]),
    "mfvgp_scg_host": (c_int, [c_void_p, _P, c_int, c_double, c_double, _P, _P, _P]),
    "mfvgp_comm_unique_id": (c_int, [_P]),
    "mfvgp_comm_init": (c_int, [c_void_p, c_int, c_int, _P]),
    "mfvgp_launch_count": (c_int64, [c_void_p]),
    "mfvgp_reconstruct": (c_int, [c_int, c_int, c_int, _P, _P, c_int64, _P, c_int64, c_int, _P, _P, _P,
                                  c_void_p]),
    "mfvgp_reconstruct_host": (c_int, [c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P]),
    "mfvgp_initial_state": (c_int, [c_int, c_int, c_int64, _P, _P, _P, c_double, c_double, _P,
                                    c_void_p]),
    "mfvgp_l96_trajectory": (c_int, [c_int, c_double, c_double, c_int, c_int, c_double, _P, _P,
                                     POINTER(c_int32), c_void_p]),
    "mfvgp_small_trajectory": (c_int, [c_int, _P, c_int, c_double, c_int, c_int, c_double, c_int,
                                       _P, _P, _P, _P, c_void_p]),
    "mfvgp_describe": (c_int, [c_void_p, c_char_p, c_int]),
    "mfvgp_refine_info": (c_int, [c_void_p, _P]),
    "mfvgp_timing": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_int64)]),
    "mfvgp_fp64_peak": (c_int, [c_int, POINTER(c_double)]),
}

_lib = None


class MfvgpError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MFVGP_LIB", LIB_PATH))
    if not path.is_file():
        raise MfvgpError(f"{path} not found: build it with `make` (or "
                         f"`python -c 'import __graft_entry__ as g; g.build()'`). "
                         f"meanfieldvargp_b200 has no CPU fallback.")
    lib = ctypes.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mfvgp_last_error().decode("utf-8", "replace")
        raise MfvgpError(f"{what} failed (code {rc}): {msg}")


def require_device():
    lib = load()
    if lib.mfvgp_device_count() < 1:
        raise MfvgpError("no CUDA device visible: the free-energy path runs "
                         "only on the GPU (no CPU fallback)")
    return lib


class _PinnedBlock(object):
    """Owner of one page-locked block of the library's pool (returned on collection)."""
    __slots__ = ("ptr", "lib")

    def __init__(self, lib, nbytes):
        self.lib = lib
        self.ptr = lib.mfvgp_host_alloc(nbytes)
        if not self.ptr:
            raise MfvgpError("mfvgp_host_alloc failed: "
                             + lib.mfvgp_last_error().decode("utf-8", "replace"))

    def __del__(self):
        try:
            self.lib.mfvgp_host_free(self.ptr)
        except Exception:
            pass


_ctype_cache = {}


def pinned_empty(n):
    """float64 numpy array of n entries in page-locked memory (recycled by the library)."""
    import numpy as np
    lib = load()
    blk = _PinnedBlock(lib, 8 * n)
    ct = _ctype_cache.get(n)
    if ct is None:
        ct = _ctype_cache[n] = c_double * n
    buf = ct.from_address(blk.ptr)
    buf._owner = blk                    # the block lives as long as any view of buf
    return np.frombuffer(buf, dtype=np.float64)
