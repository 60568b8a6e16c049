"""oracle/ref.py -- TEST INFRASTRUCTURE, not product code.

ctypes binding to oracle/_ref/libmg_ref.so, the reference's own NOCUDA_TESI solver compiled
unmodified (see build_ref.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product package (pde_multigrid_b200) never does.

    RefMG(dim, dtype, corrected=False, n=..., range=..., A=..., alfa=...)
mirrors the public methods of the reference classes MultiGrid{1,2,3}D
(N3/MultiGrid3D.h:6-33, N2/MultiGrid2D.h:6-37, N1/MultiGrid1D.h:6-31) on numpy arrays laid out
exactly like the reference (dense, x fastest: idx = x + y*sx + z*sx*sy).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libmg_ref.so")
_lib = None


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmg_ref.so missing: run `python oracle/build_ref.py` "
                               "in a container that has /root/reference")
        _lib = ctypes.CDLL(SO_PATH)
    return _lib


def norms(r):
    """Residual-norm definition of SURVEY.md 8c: l2 = sqrt(sum r^2) accumulated in fp64 in index
    order, linf = max |r|."""
    r64 = np.asarray(r, dtype=np.float64).ravel()
    return float(np.sqrt(np.cumsum(r64 * r64)[-1])) if r64.size else 0.0, float(np.max(np.abs(r64)))


class RefMG:
    def __init__(self, dim, dtype, corrected=False, n=33, range=None, A=(-1.0, -2.0, 0.0, -3.0), alfa=2, opt="O2", shape=None):
        """opt="O0": the as-shipped build (no compiler flags), available for 3D double corrected only (timing).
        shape=(nx, ny, nz): a non-cubic 3D grid -- the -DNDEBUG variants of the reference (build_ref.py: only the asserts at
        N3/Grid3D.cpp:10-11 stand between the reference and such grids)."""
        self.dim = dim
        self.np_dtype = np.dtype(dtype)
        assert self.np_dtype in (np.dtype(np.float32), np.dtype(np.float64))
        prec = "f32" if self.np_dtype == np.dtype(np.float32) else "f64"
        if corrected and dim == 2:
            corrected = False  # the 2D residual has no defect
        self.prefix = "ref%dd_%s%s%s%s" % (dim, prec, "c" if corrected else "", "" if opt == "O2" else opt, "x" if shape is not None else "")
        assert shape is None or dim == 3
        self.L = lib()
        if not hasattr(self.L, self.prefix + "_create"):
            raise RuntimeError("oracle/_ref has no variant %s" % self.prefix)
        self.creal_p = ctypes.POINTER(ctypes.c_float if prec == "f32" else ctypes.c_double)
        if range is None:
            range = [0.0, 1.0] * dim
        rng = (ctypes.c_double * (2 * dim))(*[float(x) for x in range])
        create = self._fn("create", ctypes.c_void_p)
        if dim == 2:
            a4 = (ctypes.c_double * 4)(*[float(x) for x in A])
            self.h = create(ctypes.c_int(n), rng, a4, ctypes.c_int(int(alfa)))
        elif shape is not None:
            create = self._fn("create_xyz", ctypes.c_void_p)
            self.h = create(ctypes.c_int(int(shape[0])), ctypes.c_int(int(shape[1])), ctypes.c_int(int(shape[2])), rng)
        else:
            self.h = create(ctypes.c_int(n), rng)
        self.h = ctypes.c_void_p(self.h)
        self.num_levels = self._fn("num_levels", ctypes.c_int)(self.h)
        self.sizes = [self._fn("level_size", ctypes.c_int)(self.h, ctypes.c_int(l)) for l in np.arange(self.num_levels)]
        self.shapes = None
        if shape is not None:  # numpy axis order (z, y, x) of the dense x-fastest layout
            self.shapes = []
            for l in np.arange(self.num_levels):
                o = (ctypes.c_int * 3)()
                self._fn("level_size_xyz")(self.h, ctypes.c_int(l), o)
                self.shapes.append((o[2], o[1], o[0]))

    def _fn(self, name, restype=None):
        f = getattr(self.L, "%s_%s" % (self.prefix, name))
        f.restype = restype
        return f

    def close(self):
        if self.h is not None:
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shape(self, l):
        return self.shapes[l] if self.shapes is not None else (self.sizes[l],) * self.dim

    def _view(self, which, l):
        p = self._fn("level_" + which, self.creal_p)(self.h, ctypes.c_int(int(l)))
        n = int(np.prod(self.shape(l)))
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.shape(l))

    def v(self, l=0):
        """numpy view (z,y,x order of axes == C order of the dense x-fastest layout)"""
        return self._view("v", l)

    def f(self, l=0):
        return self._view("f", l)

    def h_of(self, l=0):
        return self._fn("level_h", ctypes.c_double)(self.h, ctypes.c_int(int(l)))

    def _p(self, arr):
        assert arr.dtype == self.np_dtype and arr.flags["C_CONTIGUOUS"]
        return arr.ctypes.data_as(self.creal_p)

    # ---- operators (reference method names in comments) ----
    def relax(self, l, ncycles):  # Relax(Grid*, ncycles)
        self._fn("relax")(self.h, ctypes.c_int(int(l)), ctypes.c_int(int(ncycles)))

    def residual(self, l=0):  # CalculateResidual(Grid*)
        out = np.empty(self.shape(l), dtype=self.np_dtype)
        self._fn("residual")(self.h, ctypes.c_int(int(l)), self._p(out))
        return out

    def restrict(self, fine):  # Restrict(fine, fsize, coarse, csize)
        fn = fine.shape[0]
        cn = (fn - 1) // 2 + 1
        coarse = np.zeros((cn,) * self.dim, dtype=self.np_dtype)
        self._fn("restrict_")(self.h, self._p(fine), ctypes.c_int(fn), self._p(coarse), ctypes.c_int(cn))
        return coarse

    def interpolate(self, fine, coarse):  # Interpolate(fine, fsize, coarse, csize); fine modified in place
        self._fn("interpolate")(self.h, self._p(fine), ctypes.c_int(fine.shape[0]), self._p(coarse),
                                ctypes.c_int(coarse.shape[0]))
        return fine

    def apply_correction(self, fine, err):  # ApplyCorrection(fine, fsize, error, esize)
        self._fn("apply_correction")(self.h, self._p(fine), ctypes.c_int(fine.shape[0]), self._p(err),
                                     ctypes.c_int(err.shape[0]))
        return fine

    def set_to_value(self, grid, value, modify_boundaries):  # setToValue(grid, size, value, modifyBoundaries)
        self._fn("set_to_value")(self.h, self._p(grid), ctypes.c_int(grid.shape[0]), ctypes.c_double(float(value)),
                                 ctypes.c_int(1 if modify_boundaries else 0))
        return grid

    def vcycle(self, l, v1, v2):  # VCycle(gridID, v1, v2)
        self._fn("vcycle")(self.h, ctypes.c_int(int(l)), ctypes.c_int(int(v1)), ctypes.c_int(int(v2)))

    def fmg(self, l, v0, v1, v2):  # FullMultiGridVCycle(gridID, v0, v1, v2)
        self._fn("fmg")(self.h, ctypes.c_int(int(l)), ctypes.c_int(int(v0)), ctypes.c_int(int(v1)),
                        ctypes.c_int(int(v2)))

    def residual_norms(self, l=0):
        return norms(self.residual(l))
