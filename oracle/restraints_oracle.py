"""CPU ORACLE (test infrastructure, NOT the product): ctypes face of
restraints_oracle.c plus the SplineFunc end-knot rule.  PARITY UNPINNED against
PyRosetta (see the C file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TYPES = ("dist", "omega", "theta", "phi")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.trxo_dihedral.restype = C.c_double
        _LIB.trxo_angle.restype = C.c_double
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def apply_end_rule(x, y, bin_size, rule="H1"):
    """SplineFunc knots from the listed x/y [ROSETTA-RECALL, SURVEY 8a row 9].
    H1: extra knots (x_1-bin_size, y_1) and (x_n+bin_size, y_n); H2: none."""
    x = np.asarray(x, dtype=np.float64)
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    if rule == "H2":
        return x.copy(), y.copy()
    assert rule == "H1"
    xx = np.concatenate([[x[0] - bin_size], x, [x[-1] + bin_size]])
    yy = np.concatenate([y[:, :1], y, y[:, -1:]], axis=1)
    return xx, yy


def spline_fit(x, y):
    """Clamped (zero end slope) second derivatives, one row per restraint."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.atleast_2d(y), dtype=np.float64)
    y2 = np.empty_like(y)
    L = lib()
    for r in range(y.shape[0]):
        L.trxo_spline_fit(C.c_int(len(x)), _p(x), _p(y[r]), C.c_double(0.0), C.c_double(0.0), _p(y2[r]))
    return y2


def splinefunc(x, y, y2, xq):
    f, df = C.c_double(), C.c_double()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    y2 = np.ascontiguousarray(y2, dtype=np.float64)
    lib().trxo_splinefunc(C.c_int(len(x)), _p(x), _p(y), _p(y2), C.c_double(float(xq)), C.byref(f), C.byref(df))
    return f.value, df.value


class RestraintSetOracle:
    """Active restraints of one target: knots after the end rule + fitted y2."""

    def __init__(self, rst, sel=None, rule="H1"):
        self.sets = {}
        # 'AtomPair CA a CA b' of the af2 variant (utils_ros.py:191): distance-only sets; the C code scores
        # the CB slot, so CA is presented there and the gradient moved back (energy_grad only)
        self.dist_atom = rst.get("dist", {}).get("atom", "CB")
        assert self.dist_atom == "CB" or set(rst) == {"dist"}
        for name in TYPES:
            if name not in rst:
                continue
            rec = rst[name]
            m = np.ones(len(rec["a"]), dtype=bool) if sel is None else sel[name]
            x, y = apply_end_rule(rec["x"], rec["y"][m], rec["bin_size"], rule)
            y = np.ascontiguousarray(y.reshape(int(m.sum()), len(x)))
            self.sets[name] = dict(a=np.ascontiguousarray(rec["a"][m], dtype=np.int32),
                                   b=np.ascontiguousarray(rec["b"][m], dtype=np.int32),
                                   x=np.ascontiguousarray(x), y=y,
                                   y2=spline_fit(x, y) if len(y) else y.copy())

    def _set_args(self):
        args = []
        empty_i = np.zeros(1, dtype=np.int32)
        empty_d = np.zeros(1, dtype=np.float64)
        self._keep = (empty_i, empty_d)
        for name in TYPES:
            s = self.sets.get(name)
            if s is None or len(s["a"]) == 0:
                args += [C.c_int(0), _p(empty_i, C.c_int), _p(empty_i, C.c_int), C.c_int(2),
                         _p(empty_d), _p(empty_d), _p(empty_d)]
            else:
                args += [C.c_int(len(s["a"])), _p(s["a"], C.c_int), _p(s["b"], C.c_int), C.c_int(len(s["x"])),
                         _p(s["x"]), _p(s["y"]), _p(s["y2"])]
        return args

    def energy_grad_batch(self, xyz, w=(1.0, 1.0, 1.0), nthreads=1, want_grad=True):
        """xyz (N,L,3,3) -> (E (N,3), grad (N,L,3,3)); decoys split over host threads."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        N, L = xyz.shape[:2]
        w = np.ascontiguousarray(w, dtype=np.float64)
        E = np.zeros((N, 3))
        grad = np.zeros((N, L, 3, 3)) if want_grad else None
        lib().trxo_energy_grad_batch(C.c_int(nthreads), C.c_int(N), C.c_int(L), _p(xyz), *self._set_args(),
                                     _p(w), _p(E), _p(grad) if want_grad else None)
        return E, grad

    def energy_grad(self, xyz, w=(1.0, 1.0, 1.0)):
        """xyz (L,3,3) float64 [res][N,CA,CB][xyz] -> (E[3] unweighted, grad (L,3,3) of w.E)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        if self.dist_atom == "CA":
            swapped = xyz.copy()
            swapped[:, 2] = xyz[:, 1]
            self.dist_atom = "CB"
            try:
                E, g = self.energy_grad(swapped, w)
            finally:
                self.dist_atom = "CA"
            grad = np.zeros_like(g)
            grad[:, 1] = g[:, 2]
            return E, grad
        L = xyz.shape[0]
        args = [C.c_int(L), _p(xyz)]
        empty_i = np.zeros(1, dtype=np.int32)
        empty_d = np.zeros(1, dtype=np.float64)
        for name in TYPES:
            s = self.sets.get(name)
            if s is None or len(s["a"]) == 0:
                args += [C.c_int(0), _p(empty_i, C.c_int), _p(empty_i, C.c_int), C.c_int(2),
                         _p(empty_d), _p(empty_d), _p(empty_d)]
            else:
                args += [C.c_int(len(s["a"])), _p(s["a"], C.c_int), _p(s["b"], C.c_int), C.c_int(len(s["x"])),
                         _p(s["x"]), _p(s["y"]), _p(s["y2"])]
        w = np.ascontiguousarray(w, dtype=np.float64)
        E = np.zeros(3)
        grad = np.zeros((L, 3, 3))
        lib().trxo_energy_grad_flat(*args, _p(w), _p(E), _p(grad))
        return E, grad
