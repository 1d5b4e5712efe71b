"""ctypes binding of libtrx2dyn.so (include/trx2dyn.h).  No CPU fallback: loading or
creating a context without the CUDA library / a B200 raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

F64, F32 = 64, 32
_DT = {F64: np.float64, F32: np.float32}
_lib = None


class TrxError(RuntimeError):
    pass


class RstSet(C.Structure):
    _fields_ = [("n", C.c_int), ("a", C.POINTER(C.c_int32)), ("b", C.POINTER(C.c_int32)),
                ("K", C.c_int), ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double))]


def lib():
    """Loads (building if sources are newer) the in-tree CUDA library."""
    global _lib
    if _lib is None:
        path = os.environ.get("TRX2DYN_LIB") or _build.LIB      # TRX2DYN_LIB: load a specific build (A/B tests)
        if path == _build.LIB and _build.needs_build():
            if not os.path.exists(_build.NVCC):
                if not os.path.exists(path):
                    raise TrxError("libtrx2dyn.so is missing and nvcc is not available; there is no CPU fallback")
            else:
                _build.build_library()
        L = C.CDLL(path)
        L.trx_last_error.restype = C.c_char_p
        L.trx_ctx_launch_count.restype = C.c_longlong
        L.trx_ctx_launch_count.argtypes = [C.c_void_p]
        for fn in (L.trx_fold_run_queue, L.trx_fold_mc_queue, L.trx_fold_k1_evals, L.trx_fold_status, L.trx_dyn_create, L.trx_dyn_step, L.trx_dyn_get, L.trx_dyn_destroy):
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise TrxError("libtrx2dyn: %s (status %d)" % (lib().trx_last_error().decode(), rc))


def _ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class Context:
    def __init__(self, device=0, stream=None):
        import weakref
        self._h = C.c_void_p()
        check(lib().trx_ctx_create(C.c_int(device), C.c_void_p(stream or 0), C.byref(self._h)))
        self.device = device
        self._children = weakref.WeakSet()   # tables, fold batches, dynamics states: they hold pointers into the context

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        if self._h:
            for child in list(self._children):   # a handle that outlives its context would dangle (its destroy reads the context)
                child.close()
            lib().trx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().trx_ctx_sync(self._h))

    def set_timing(self, on=True):
        """on: False / True (every kernel) / 2 (only the restraint kernel and the whole fold)."""
        check(lib().trx_ctx_set_timing(self._h, C.c_int(int(on))))

    def reset_timing(self):
        check(lib().trx_ctx_reset_timing(self._h))

    def timing(self, name):
        ms, n = C.c_double(), C.c_longlong()
        check(lib().trx_ctx_get_timing(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launch_count(self):
        return int(lib().trx_ctx_launch_count(self._h))


class Tables:
    """Device-resident spline tables of one target (one restraint set)."""

    def __init__(self, ctx, L, active, dist_atom="CB"):
        """active: {type: (a, b, x, y)} as tables.active_restraints returns.  dist_atom: 'CB', or 'CA' for
        the distance-only tables of the -r af2 variant (also taken from active['dist_atom'] if present)."""
        self.ctx, self.L = ctx, L
        dist_atom = active.get("dist_atom", dist_atom)
        sets = (RstSet * 4)()
        keep = []
        for t, name in enumerate(("dist", "omega", "theta", "phi")):
            if name in active and len(active[name][0]):
                a, b, x, y = active[name]
                a = np.ascontiguousarray(a, dtype=np.int32)
                b = np.ascontiguousarray(b, dtype=np.int32)
                x = np.ascontiguousarray(x, dtype=np.float64)
                y = np.ascontiguousarray(y, dtype=np.float64)
                if y.shape != (len(a), len(x)):
                    raise ValueError("%s: y must be (n, K)" % name)
                keep += [a, b, x, y]
                sets[t] = RstSet(len(a), _ptr(a, C.c_int32), _ptr(b, C.c_int32), len(x), _ptr(x, C.c_double), _ptr(y, C.c_double))
            else:
                sets[t] = RstSet(0, None, None, 0, None, None)
        self._h = C.c_void_p()
        check(lib().trx_tables_create(ctx._h, C.c_int(L), sets, C.byref(self._h)))
        ctx._adopt(self)
        self.counts = [int(s.n) for s in sets]
        self.K = [int(s.K) for s in sets]
        if dist_atom not in ("CA", "CB"):
            raise ValueError("dist_atom must be 'CA' or 'CB'")
        if dist_atom == "CA":
            check(lib().trx_tables_set_dist_atom(self._h, C.c_int(1)))

    def close(self):
        if self._h:
            lib().trx_tables_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        L, tiles = C.c_int(), C.c_int()
        counts = (C.c_int * 4)()
        check(lib().trx_tables_info(self._h, C.byref(L), counts, C.byref(tiles)))
        return dict(L=L.value, counts=list(counts), tiles=tiles.value)

    def y2(self, t):
        out = np.zeros((self.counts[t], self.K[t]))
        if out.size:
            check(lib().trx_tables_get_y2(self._h, C.c_int(t), _ptr(out, C.c_double)))
        return out

    def energy_grad(self, xyz, w=(1.0, 1.0, 1.0), precision=F64, want_grad=True):
        """xyz (N, L, 3, 3) host array [decoy][res][N,CA,CB][xyz] -> (E (N,3) float64, grad like xyz)."""
        dt = _DT[precision]
        xyz = np.ascontiguousarray(xyz, dtype=dt)
        N = xyz.shape[0]
        if xyz.shape != (N, self.L, 3, 3):
            raise ValueError("xyz must be (N, %d, 3, 3)" % self.L)
        w = np.ascontiguousarray(w, dtype=np.float64)
        E = np.zeros((N, 3))
        grad = np.zeros_like(xyz) if want_grad else None
        check(lib().trx_energy_grad(self.ctx._h, self._h, C.c_int(N), C.c_int(precision), C.c_void_p(xyz.ctypes.data),
                                    _ptr(w, C.c_double), _ptr(E, C.c_double),
                                    C.c_void_p(grad.ctypes.data) if want_grad else None))
        return E, grad

    def energy_grad_device(self, N, d_xyz, d_E, d_grad, w=(1.0, 1.0, 1.0), precision=F32):
        """Device pointers (ints), grouped layout; asynchronous on the context's stream."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        check(lib().trx_energy_grad_device(self.ctx._h, self._h, C.c_int(N), C.c_int(precision), C.c_void_p(d_xyz),
                                           _ptr(w, C.c_double), C.c_void_p(d_E), C.c_void_p(d_grad or 0)))


def padded_length(L):
    return int(lib().trx_padded_length(C.c_int(L)))


def to_grouped(ctx, N, L, n_atoms, precision, d_nat, d_grp):
    check(lib().trx_to_grouped(ctx._h, C.c_int(N), C.c_int(L), C.c_int(n_atoms), C.c_int(precision), C.c_void_p(d_nat), C.c_void_p(d_grp)))


def from_grouped(ctx, N, L, n_atoms, precision, d_grp, d_nat):
    check(lib().trx_from_grouped(ctx._h, C.c_int(N), C.c_int(L), C.c_int(n_atoms), C.c_int(precision), C.c_void_p(d_grp), C.c_void_p(d_nat)))


NTERM = 8


def _weights(w):
    """NTERM weights; the 7-term form (no H-bond weight) is padded with 0."""
    w = np.asarray(w, dtype=np.float64)
    if w.shape == (NTERM - 1,):
        w = np.concatenate([w, [0.0]])
    if w.shape != (NTERM,):
        raise ValueError("weights must have %d entries" % NTERM)
    return np.ascontiguousarray(w)


class Run(C.Structure):
    """trx_run: one MinMover.apply of the schedule."""
    _fields_ = [("w", C.c_double * NTERM), ("max_iter", C.c_int), ("tol", C.c_double),
                ("clash_check", C.c_int), ("clash_thr", C.c_double), ("skip_to", C.c_int), ("cartesian", C.c_int)]


class FoldBatch:
    """N decoys of one target folded together on one GPU (trx_fold_*)."""

    def __init__(self, ctx, tabs, ndecoys, aa, runs, lbfgs_m=20):
        self.ctx, self.tabs = ctx, list(tabs)
        self.ndecoys = [int(n) for n in ndecoys]
        self.N, self.L = sum(self.ndecoys), self.tabs[0].L
        self.nruns = len(runs)
        aa = np.ascontiguousarray(aa, dtype=np.int32)
        arr_t = (C.c_void_p * len(self.tabs))(*[t._h for t in self.tabs])
        arr_n = (C.c_int * len(self.tabs))(*self.ndecoys)
        arr_r = (Run * len(runs))(*runs)
        self._h = C.c_void_p()
        check(lib().trx_fold_create(ctx._h, C.c_int(len(self.tabs)), arr_t, arr_n, _ptr(aa, C.c_int32), arr_r,
                                    C.c_int(len(runs)), C.c_int(lbfgs_m), C.byref(self._h)))
        ctx._adopt(self)

    def close(self):
        if self._h:
            lib().trx_fold_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, tors, max_rounds=20000, check_every=16, want_xyz=True):
        """tors (N,L,3) float32 radians -> dict(tors, xyz (N,L,5,3) [N,CA,CB,C,O], terms (N,8), evals, iters, rounds)."""
        tors = np.ascontiguousarray(tors, dtype=np.float32).copy()
        if tors.shape != (self.N, self.L, 3):
            raise ValueError("tors must be (%d, %d, 3)" % (self.N, self.L))
        xyz = np.zeros((self.N, self.L, 5, 3), dtype=np.float32) if want_xyz else None
        terms = np.zeros((self.N, NTERM))
        stats = np.zeros((self.N, 2), dtype=np.int64)
        rounds = C.c_int()
        self._last_n = self.N
        check(lib().trx_fold_run(self._h, _ptr(tors, C.c_float), _ptr(xyz, C.c_float) if want_xyz else None,
                                 _ptr(terms, C.c_double), _ptr(stats, C.c_longlong), C.c_int(max_rounds),
                                 C.c_int(check_every), C.byref(rounds)))
        return dict(tors=tors, xyz=xyz, terms=terms, evals=stats[:, 0], iters=stats[:, 1], rounds=rounds.value)

    def _nq(self, tors, nq):
        nq = self.ndecoys if nq is None else [int(n) for n in nq]
        if len(nq) != len(self.ndecoys):
            raise ValueError("nq must give a decoy count per table block")
        tors = np.ascontiguousarray(tors, dtype=np.float32).copy()
        if tors.shape != (sum(nq), self.L, 3):
            raise ValueError("tors must be (%d, %d, 3)" % (sum(nq), self.L))
        return tors, nq, (C.c_int * len(nq))(*nq)

    def run_queue(self, tors, nq, max_rounds=1 << 30, check_every=16, want_xyz=True):
        """Continuous batching (trx_fold_run_queue): folds nq[t] decoys against table block t through the
        batch's positions, refilling a position as soon as its decoy has left the schedule segment in
        progress.  tors (sum nq, L, 3), the decoys of block 0 first.  Same dict as run()."""
        tors, nq, arr = self._nq(tors, nq)
        n = self._last_n = sum(nq)
        xyz = np.zeros((n, self.L, 5, 3), dtype=np.float32) if want_xyz else None
        terms = np.zeros((n, NTERM))
        stats = np.zeros((n, 2), dtype=np.int64)
        rounds = C.c_int()
        check(lib().trx_fold_run_queue(self._h, arr, _ptr(tors, C.c_float), _ptr(xyz, C.c_float) if want_xyz else None,
                                       _ptr(terms, C.c_double), _ptr(stats, C.c_longlong), C.c_int(max_rounds),
                                       C.c_int(check_every), C.byref(rounds)))
        return dict(tors=tors, xyz=xyz, terms=terms, evals=stats[:, 0], iters=stats[:, 1], rounds=rounds.value)

    def status(self, n=None):
        """TRX_DECOY_* bits of the decoys of the last run*/run_mc call (0 = clean; 1 non-finite energy,
        2 line search failed, 4 round budget ran out)."""
        n = self._last_n if n is None else n
        out = np.zeros(n, dtype=np.int32)
        check(lib().trx_fold_status(self._h, _ptr(out, C.c_int32), C.c_int(n)))
        return out

    def k1_evals(self):
        """Decoy evaluations the restraint kernel made in the last run*/run_mc call, per table block."""
        out = (C.c_longlong * len(self.ndecoys))()
        check(lib().trx_fold_k1_evals(self._h, out))
        return [int(v) for v in out]

    def run_mc(self, tors, cycles, kT=2.0, block=(3, 9), sigma_deg=20.0, seed=0, id_offset=0, max_rounds=1 << 30,
               check_every=16, nq=None):
        """Fold, then `cycles` Monte-Carlo cycles (perturb / re-minimise with the schedule's LAST run /
        Metropolis) on device.  nq: decoys per table block (continuous batching), default one per position.
        Returns the dict of run() plus 'accepted' (N,)."""
        tors, nq, arr = self._nq(tors, nq)
        n = self._last_n = sum(nq)
        xyz = np.zeros((n, self.L, 5, 3), dtype=np.float32)
        terms = np.zeros((n, NTERM))
        stats = np.zeros((n, 3), dtype=np.int64)
        rounds = C.c_int()
        check(lib().trx_fold_mc_queue(self._h, arr, _ptr(tors, C.c_float), _ptr(xyz, C.c_float), _ptr(terms, C.c_double),
                                      _ptr(stats, C.c_longlong), C.c_int(self.nruns - 1), C.c_int(cycles), C.c_double(kT),
                                      C.c_int(block[0]), C.c_int(block[1]), C.c_double(sigma_deg), C.c_ulonglong(seed),
                                      C.c_ulonglong(id_offset), C.c_int(max_rounds), C.c_int(check_every), C.byref(rounds)))
        return dict(tors=tors, xyz=xyz, terms=terms, evals=stats[:, 0], iters=stats[:, 1], accepted=stats[:, 2],
                    rounds=rounds.value)

    def eval(self, tors, w):
        """Single evaluation -> (total (N,), terms (N,8), gtors (N,L,3), xyz (N,L,5,3))."""
        tors = np.ascontiguousarray(tors, dtype=np.float32)
        w = _weights(w)
        total, terms = np.zeros(self.N), np.zeros((self.N, NTERM))
        gt = np.zeros((self.N, self.L, 3), dtype=np.float32)
        xyz = np.zeros((self.N, self.L, 5, 3), dtype=np.float32)
        check(lib().trx_fold_eval(self._h, _ptr(tors, C.c_float), _ptr(w, C.c_double), _ptr(total, C.c_double),
                                  _ptr(terms, C.c_double), _ptr(gt, C.c_float), _ptr(xyz, C.c_float)))
        return total, terms, gt, xyz

    def eval_cart(self, xyz, w):
        """Single Cartesian-mode evaluation: xyz (N,L,5,3) -> (total (N,), terms (N,8), grad (N,L,5,3), tors (N,L,3))."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        if xyz.shape != (self.N, self.L, 5, 3):
            raise ValueError("xyz must be (%d, %d, 5, 3)" % (self.N, self.L))
        w = _weights(w)
        total, terms = np.zeros(self.N), np.zeros((self.N, NTERM))
        grad = np.zeros((self.N, self.L, 5, 3), dtype=np.float32)
        tors = np.zeros((self.N, self.L, 3), dtype=np.float32)
        check(lib().trx_fold_eval_cart(self._h, _ptr(xyz, C.c_float), _ptr(w, C.c_double), _ptr(total, C.c_double),
                                       _ptr(terms, C.c_double), _ptr(grad, C.c_float), _ptr(tors, C.c_float)))
        return total, terms, grad, tors


class DynState:
    """Device-resident distograms of one chain of the outer dynamics loop (trx_dyn_*): step() applies one decoy."""

    def __init__(self, ctx, npz, angle=True):
        self.ctx = ctx
        self.L = int(np.asarray(npz["dist"]).shape[0])
        self.angle = bool(angle) and all(k in npz for k in ("omega", "theta", "phi"))
        arrs = [np.ascontiguousarray(npz["dist"], dtype=np.float32)]
        if self.angle:
            arrs += [np.ascontiguousarray(npz[k], dtype=np.float32) for k in ("omega", "theta", "phi")]
        ptrs = [_ptr(a, C.c_float) for a in arrs] + [None] * (4 - len(arrs))
        self._h = C.c_void_p()
        check(lib().trx_dyn_create(ctx._h, C.c_int(self.L), *ptrs, C.byref(self._h)))
        ctx._adopt(self)

    def close(self):
        if self._h:
            lib().trx_dyn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def gaussian_taps(sigma=1.0, truncate=4.0):
        """The kernel scipy.ndimage.gaussian_filter1d builds (_gaussian_kernel1d, order 0)."""
        lw = int(truncate * float(sigma) + 0.5)
        if lw != 4:
            raise ValueError("the device filter has 9 taps (sigma 1)")
        x = np.arange(-lw, lw + 1)
        p = np.exp(-0.5 / (sigma * sigma) * x ** 2)
        return np.ascontiguousarray(p / p.sum(), dtype=np.float64)

    def step(self, n, ca, c, cb=None, seq=None, sigma=1.0):
        """One decoy: updates the maps in place, returns max |tmp_new - tmp_old|."""
        n, ca, c = (np.ascontiguousarray(a, dtype=np.float64) for a in (n, ca, c))
        use = np.zeros(self.L, dtype=np.uint8)
        cbv = np.zeros((self.L, 3))
        if cb is not None and seq is not None:
            cb = np.asarray(cb, dtype=np.float64)
            use = (np.array([s != "G" for s in seq]) & ~np.isnan(cb).any(axis=1)).astype(np.uint8)
            cbv = np.ascontiguousarray(np.nan_to_num(cb))
        w = self.gaussian_taps(sigma)
        chg = C.c_double()
        check(lib().trx_dyn_step(self._h, _ptr(n, C.c_double), _ptr(ca, C.c_double), _ptr(c, C.c_double), _ptr(cbv, C.c_double),
                                 _ptr(use, C.c_ubyte), _ptr(w, C.c_double), C.byref(chg)))
        return chg.value

    def get(self, want_bins=False):
        """Current maps as a dict like the reference's npz (dist, [omega, theta, phi,] tmp[, bins])."""
        L = self.L
        out = {"dist": np.empty((L, L, 37), dtype=np.float32), "tmp": np.empty((L, L, 37), dtype=np.float32)}
        if self.angle:
            out.update(omega=np.empty((L, L, 25), dtype=np.float32), theta=np.empty((L, L, 25), dtype=np.float32),
                       phi=np.empty((L, L, 13), dtype=np.float32))
        bins = np.empty((4, L, L), dtype=np.int32) if want_bins else None
        check(lib().trx_dyn_get(self._h, _ptr(out["dist"], C.c_float),
                                _ptr(out["omega"], C.c_float) if self.angle else None,
                                _ptr(out["theta"], C.c_float) if self.angle else None,
                                _ptr(out["phi"], C.c_float) if self.angle else None,
                                _ptr(out["tmp"], C.c_float), _ptr(bins, C.c_int32) if want_bins else None))
        if want_bins:
            out["bins"] = bins
        return out
