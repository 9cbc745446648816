"""ctypes binding of oracle/c/liboracle.so (the plain-C, multi-threaded restatement of the
reference's five FEM kernels).  Test / CPU-baseline infrastructure only."""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "c"
KINDS = {"snh": 0, "arap": 1, "muscle": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        L = ctypes.CDLL(str(so))
        L.oracle_fun.restype = ctypes.c_double
        L.oracle_hess_quad.restype = ctypes.c_double
        _lib = L
    return _lib


class CPotential:
    """Same operator surface as oracle.fem.FemPotential, fp64, multi-threaded."""

    def __init__(self, kind: str, cells, dhdX, dV, mu, lambda_=None, activation=None):
        self.kind = KINDS[kind]
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.dhdX, self.dV, self.mu, self.la, self.act = c(dhdX), c(dV), c(mu), c(lambda_), c(activation)
        self.T = self.cells.shape[0]

    def _args(self):
        P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        return (self.kind, ctypes.c_int64(self.T), P(self.cells), P(self.dhdX), P(self.dV), P(self.mu), P(self.la),
                P(self.act))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    def fun(self, u, output):
        u = self._c(u)
        output[0] += lib().oracle_fun(*self._args(), self._p(u))

    def grad(self, u, output):
        u = self._c(u)
        lib().oracle_grad(*self._args(), self._p(u), self._p(output))

    def hess_diag(self, u, output):
        u = self._c(u)
        lib().oracle_hess_diag(*self._args(), self._p(u), self._p(output))

    def hess_prod(self, u, p, output):
        u, p = self._c(u), self._c(p)
        lib().oracle_hess_prod(*self._args(), self._p(u), self._p(p), self._p(output))

    def hess_quad(self, u, p, output):
        u, p = self._c(u), self._c(p)
        output[0] += lib().oracle_hess_quad(*self._args(), self._p(u), self._p(p))


def num_threads() -> int:
    return int(lib().oracle_num_threads())
