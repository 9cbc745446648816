"""ctypes binding of oracle/c/liboracle.so (the plain-C, multi-threaded restatement of the
reference's five FEM kernels).  Test / CPU-baseline infrastructure only."""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "c"
KINDS = {"snh": 0, "arap": 1, "muscle": 2}
_libs = {}


def lib(dtype=np.float64):
    """liboracle.so (fp64) or liboracle_f32.so (the same source compiled with -DORACLE_F32)."""
    key = np.dtype(dtype).itemsize
    if key not in _libs:
        so = HERE / ("liboracle.so" if key == 8 else "liboracle_f32.so")
        if not so.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        L = ctypes.CDLL(str(so))
        L.oracle_fun.restype = ctypes.c_double
        L.oracle_hess_quad.restype = ctypes.c_double
        _libs[key] = L
    return _libs[key]


class CPotential:
    """Same operator surface as oracle.fem.FemPotential, multi-threaded; ``dtype`` float64 (default) or float32
    (inputs and outputs of the operators are then float32 arrays)."""

    def __init__(self, kind: str, cells, dhdX, dV, mu, lambda_=None, activation=None, dtype=np.float64):
        self.kind = KINDS[kind]
        self.dtype = np.dtype(dtype)
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=self.dtype)  # noqa: E731
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

    def _c(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    def _out(self, output):
        if output.dtype != self.dtype or not output.flags.c_contiguous:
            raise TypeError(f"output must be a contiguous {self.dtype} array")
        return self._p(output)

    def fun(self, u, output):
        u = self._c(u)
        output[0] += lib(self.dtype).oracle_fun(*self._args(), self._p(u))

    def grad(self, u, output):
        u = self._c(u)
        lib(self.dtype).oracle_grad(*self._args(), self._p(u), self._out(output))

    def hess_diag(self, u, output):
        u = self._c(u)
        lib(self.dtype).oracle_hess_diag(*self._args(), self._p(u), self._out(output))

    def hess_prod(self, u, p, output):
        u, p = self._c(u), self._c(p)
        lib(self.dtype).oracle_hess_prod(*self._args(), self._p(u), self._p(p), self._out(output))

    def hess_quad(self, u, p, output):
        u, p = self._c(u), self._c(p)
        output[0] += lib(self.dtype).oracle_hess_quad(*self._args(), self._p(u), self._p(p))


def num_threads() -> int:
    return int(lib().oracle_num_threads())
