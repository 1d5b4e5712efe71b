"""CPU ORACLE (test infrastructure, NOT the product): ctypes face of fold_oracle.c.
PARITY UNPINNED against PyRosetta (see the C file's header)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .restraints_oracle import lib, TYPES

NTERM = 8
AA_ORDER = "ARNDCQEGHILKMFPSTWYV"


class Run(C.Structure):
    _fields_ = [("w", C.c_double * NTERM), ("max_iter", C.c_int), ("tol", C.c_double),
                ("clash_check", C.c_int), ("clash_thr", C.c_double), ("skip_to", C.c_int), ("cartesian", C.c_int)]


def _weights(w):
    """NTERM weights; the 7-term form of earlier callers (no H-bond weight) is padded with 0."""
    w = np.asarray(w, dtype=np.float64)
    if w.shape == (NTERM - 1,):
        w = np.concatenate([w, [0.0]])
    assert w.shape == (NTERM,)
    return np.ascontiguousarray(w)


def aa_index(seq, gly_to_ala=True):
    """folding.py:112-115 mutates G->A for the centroid stage."""
    idx = np.array([AA_ORDER.index(c) if c in AA_ORDER else 0 for c in seq], dtype=np.int32)
    if gly_to_ala:
        idx[idx == AA_ORDER.index("G")] = 0
    return idx


def reference_schedule(cartesian=True):
    """folding.py:74-104,118-119,164-171 (mode 2) with data/*.wts.  Term order: apc, dih, ang, vdw, rama,
    omega, cart_bonded, backbone H-bond (cen_hb in the centroid stages, hbond_sr_bb / hbond_lr_bb in the
    Cartesian one: a stated approximation, include/trx_centroid_model.h)."""
    def run(w, it, clash=False, skip_to=0, cart=False):
        r = Run()
        r.w[:] = w
        r.max_iter, r.tol = it, 1e-4
        r.clash_check, r.clash_thr, r.skip_to, r.cartesian = int(clash), 10.0, skip_to, int(cart)
        return r
    sf = [5, 4, 4, 1, 1, 0.5, 0, 5]              # ... cen_hb 5
    sf1 = [3, 1, 1, 3, 1, 0.5, 0, 5]
    sf_vdw = [0, 0, 0, 1, 1, 0, 0, 0]
    sf_cart = [5, 4, 4, 0.5, 1, 0.5, 0.1, 3]     # ... hbond_sr_bb 3, hbond_lr_bb 3 (one function here, one weight)
    runs = [run(sf_vdw, 500, True, 5) for _ in range(5)]          # remove_clash(sf_vdw, min_mover_vdw)
    runs += [run(sf, 1000) for _ in range(3)]                     # RepeatMover(min_mover, 3)
    if cartesian:
        runs += [run(sf_cart, 1000, cart=True)]                   # min_mover_cart
    end = len(runs) + 5
    runs += [run(sf1, 1000, True, end) for _ in range(5)]         # remove_clash(sf_vdw, min_mover1)
    return runs


class FoldOracle:
    def __init__(self, rs, seq):
        """rs: RestraintSetOracle (or None for no restraints); seq: one-letter string."""
        self.L = len(seq)
        self.aa = aa_index(seq)
        self._keep = []
        n, K = (C.c_int * 4)(), (C.c_int * 4)()
        pa, pb = (C.POINTER(C.c_int) * 4)(), (C.POINTER(C.c_int) * 4)()
        px, py, py2 = (C.POINTER(C.c_double) * 4)(), (C.POINTER(C.c_double) * 4)(), (C.POINTER(C.c_double) * 4)()
        for t, name in enumerate(TYPES):
            s = rs.sets.get(name) if rs is not None else None
            if s is None or len(s["a"]) == 0:
                n[t], K[t] = 0, 0
                continue
            n[t], K[t] = len(s["a"]), len(s["x"])
            pa[t] = s["a"].ctypes.data_as(C.POINTER(C.c_int))
            pb[t] = s["b"].ctypes.data_as(C.POINTER(C.c_int))
            px[t] = s["x"].ctypes.data_as(C.POINTER(C.c_double))
            py[t] = s["y"].ctypes.data_as(C.POINTER(C.c_double))
            py2[t] = s["y2"].ctypes.data_as(C.POINTER(C.c_double))
            self._keep.append(s)
        self._args = (n, pa, pb, K, px, py, py2)
        lib().trxo_eval_flat.restype = C.c_double

    def nerf(self, tors):
        tors = np.ascontiguousarray(tors, dtype=np.float64)
        xyz = np.zeros((self.L, 5, 3))
        lib().trxo_nerf(C.c_int(self.L), tors.ctypes.data_as(C.POINTER(C.c_double)), xyz.ctypes.data_as(C.POINTER(C.c_double)))
        return xyz

    def eval(self, tors, w):
        """-> (total, terms[8], gtors (L,3), xyz (L,5,3))"""
        tors = np.ascontiguousarray(tors, dtype=np.float64)
        w = _weights(w)
        terms, gt, xyz = np.zeros(NTERM), np.zeros((self.L, 3)), np.zeros((self.L, 5, 3))
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        tot = lib().trxo_eval_flat(C.c_int(self.L), self.aa.ctypes.data_as(C.POINTER(C.c_int)), *self._args, P(w), P(tors),
                                   P(terms), P(gt), P(xyz))
        return tot, terms, gt, xyz

    def eval_cart(self, xyz, w):
        """Cartesian-mode evaluation -> (total, terms[8], grad (L,5,3))."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        w = _weights(w)
        terms, g = np.zeros(NTERM), np.zeros((self.L, 5, 3))
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        lib().trxo_eval_cart_flat.restype = C.c_double
        tot = lib().trxo_eval_cart_flat(C.c_int(self.L), self.aa.ctypes.data_as(C.POINTER(C.c_int)), *self._args, P(w), P(xyz),
                                        P(terms), P(g))
        return tot, terms, g

    def torsions(self, xyz):
        """(phi, psi, omega) read back from coordinates (L,5,3) -> (L,3)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        t = np.zeros((self.L, 3))
        lib().trxo_torsions_from_xyz(C.c_int(self.L), xyz.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)))
        return t

    def fold(self, tors0, runs, m=20, nthreads=1):
        """tors0 (N,L,3) -> dict(tors, terms (N,7), xyz (N,L,5,3), f, evals, iters)."""
        tors = np.ascontiguousarray(tors0, dtype=np.float64).copy()
        N = tors.shape[0]
        arr = (Run * len(runs))(*runs)
        terms, xyz = np.zeros((N, NTERM)), np.zeros((N, self.L, 5, 3))
        f, ev, it = np.zeros(N), np.zeros(N, dtype=np.int64), np.zeros(N, dtype=np.int64)
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        lib().trxo_fold_batch(C.c_int(nthreads), C.c_int(N), C.c_int(self.L), self.aa.ctypes.data_as(C.POINTER(C.c_int)),
                              *self._args, arr, C.c_int(len(runs)), C.c_int(m), P(tors), P(terms), P(xyz), P(f),
                              ev.ctypes.data_as(C.POINTER(C.c_longlong)), it.ctypes.data_as(C.POINTER(C.c_longlong)))
        return dict(tors=tors, terms=terms, xyz=xyz, f=f, evals=ev, iters=it)


def random_torsions(N, L, seed):
    """set_random_dihedral (utils_ros.py:656-696): residues 1..L-1 draw (phi,psi) from the
    6-state table, omega = 180; residue L keeps 180s.  Radians, (N,L,3)."""
    rng = np.random.default_rng(seed)
    states = np.array([[-140, 153], [-72, 145], [-122, 117], [-82, -14], [-61, -41], [57, 39]], dtype=np.float64)
    edges = np.array([0.135, 0.29, 0.363, 0.485, 0.982])
    r = rng.random((N, L))
    k = np.searchsorted(edges, r, side="left")   # r <= edge -> that state
    t = np.full((N, L, 3), 180.0)
    t[:, :, :2] = states[k]
    t[:, L - 1, :] = 180.0
    return np.deg2rad(t)
