"""ctypes loader for libmg_b200.so (the C-ABI library declared in include/mg_b200.h).

There is no CPU fallback: if the library has not been built, loading raises; if it is loaded on a
machine without a CUDA device, every create() call raises MGError(MG_ERR_CUDA).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MG_B200_LIB: another build of the same library (the -DMG_DEBUG_BOUNDS one, tests/test_debug_bounds.py)
SO_PATH = os.environ.get("MG_B200_LIB") or os.path.join(HERE, "libmg_b200.so")

MG_OK, MG_ERR_ARG, MG_ERR_CUDA, MG_ERR_NOMEM, MG_ERR_STATE, MG_ERR_COMM = 0, 1, 2, 3, 4, 5
MG_F32, MG_F64 = 0, 1
MG_REF_COMPAT, MG_CORRECTED = 0, 1
MG_FIELD_V, MG_FIELD_F = 0, 1
MG_SMOOTHER_AUTO, MG_SMOOTHER_COLOUR, MG_SMOOTHER_FUSED, MG_SMOOTHER_JACOBI, MG_SMOOTHER_TMA, MG_SMOOTHER_PIPE = 0, 1, 2, 3, 4, 5
MG_ARITH_EXACT, MG_ARITH_FAST = 0, 1

_lib = None


class MGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mg_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load the C-ABI library; raise loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("%s not built: run `python -m pde_multigrid_b200.build` (needs nvcc); "
                              "pde_multigrid_b200 has no CPU or PyTorch fallback" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        L.mg_last_error.restype = ctypes.c_char_p
        L.mg_version.restype = ctypes.c_char_p
        for dim in ("1d", "2d", "3d"):
            for name, rt in (("level_h", ctypes.c_double), ("stream", ctypes.c_void_p),
                             ("kernel_launches", ctypes.c_longlong),
                             ("halo_bytes", ctypes.c_longlong)):
                f = getattr(L, "mg%s_%s" % (dim, name), None)
                if f is not None:
                    f.restype = rt
        _lib = L
    return _lib


def check(status):
    if status != MG_OK:
        raise MGError(status, lib().mg_last_error().decode("utf-8", "replace"))
