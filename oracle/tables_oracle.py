"""CPU ORACLE (test infrastructure, NOT the product) -- restraint-table construction.

Plain numpy restatement of the reference's distogram -> Rosetta SPLINE restraint
recipe.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this file.  The shipped path is trx2dyn.tables (fast, vectorised) which is
tested against this one; this one is tested against the reference's own gen_rst
imported with a stub ``pyrosetta`` (tests/golden/make_golden.py).

Follows, line by line:
  /root/reference/folding/utils_ros/utils_ros.py:6-146    gen_rst
  /root/reference/folding/utils_ros/utils_ros.py:706-743  add_rst (selection only)
  /root/reference/folding/data/params.json:1-16           constants

The reference writes each table as two lines of text ('x_axis\\t%.3f...' and
'y_axis\\t%.3f...'; '%.5f' for omega) which Rosetta re-reads, so the knots the
scoring code sees are the DECIMAL-ROUNDED values.  We keep the text itself
(`text_lines`) so equality with the reference is checked on bytes.
"""
from __future__ import annotations

import numpy as np

PARAMS = {  # folding/data/params.json:1-16
    "PCUT": 0.05, "PCUT1": 0.5, "EBASE": -0.5, "EREP": [10.0, 3.0, 0.5],
    "DREP": [0.0, 2.0, 3.5], "PREP": 0.1, "SIGD": 10.0, "SIGM": 1.0,
    "MEFF": 0.0001, "DCUT": 19.5, "ALPHA": 1.57, "DSTEP": 0.5, "ASTEP": 15.0,
}


def _fmt_rows(arr, fmt):
    """'%.3f'-format every element and parse back (what Rosetta reads)."""
    s = np.char.mod(fmt, arr)
    return s, s.astype(np.float64)


def gen_rst_oracle(npz, use_orient=True, params=PARAMS):
    """Restates gen_rst (utils_ros.py:6-146).  Returns dict type -> dict with
    a, b (0-based residue indices, reference order = np.where row-major order),
    p (float32 probability), x (K,) float64 knots, y (n,K) float64 knots (both
    after decimal rounding), xs/ys the text tokens, bin_size (float from
    '%.5f' text on the restraint line)."""
    MEFF, DCUT, ALPHA = params["MEFF"], params["DCUT"], params["ALPHA"]
    EBASE, EREP, DREP = params["EBASE"], params["EREP"], params["DREP"]
    DSTEP = params["DSTEP"]
    ASTEP = np.deg2rad(params["ASTEP"])
    PCUT = 0.05  # utils_ros.py:18 (hard-coded, ignores -pd)
    out = {}

    # ---- dist (utils_ros.py:54-75): 3 repulsive knots + 32 attractive ones
    dist = npz["dist"]                                   # float32 (L,L,37)
    centres = 4.25 + DSTEP * np.arange(32)               # bin centres 4.25..19.75
    p_contact = dist[..., 5:].sum(axis=-1)               # float32 sum (:56)
    background = (centres / DCUT) ** ALPHA               # float64 (:57)
    numer = dist[..., 5:] + MEFF                         # stays float32
    denom = dist[..., 36:37] * background + 1e-6         # float32*float64 -> float64
    e_attr = EBASE - np.log(numer / denom)               # == -log(.)+EBASE (:58)
    e_first = np.where(e_attr[..., 0] > 0.0, e_attr[..., 0], 0.0)
    e_rep = e_first[..., None] + np.asarray(EREP)        # (:59)
    tab = np.concatenate([e_rep, e_attr], axis=-1)       # 35 knots (:60)
    knots = np.concatenate([np.asarray(DREP), centres])  # (:61)
    i, j = np.nonzero(p_contact > PCUT)                  # row-major order (:62)
    upper = j > i                                        # (:67)
    i, j = i[upper], j[upper]
    xs, x = _fmt_rows(knots, "%.3f")
    ys, y = _fmt_rows(tab[i, j], "%.3f")
    out["dist"] = dict(a=i, b=j, p=p_contact[i, j], x=x, y=y, xs=xs, ys=ys,
                       bin_size=float("%.5f" % 0.5))
    if not use_orient:
        return out

    def neg_log_ratio(arr):
        # -log((p+MEFF)/(p_last+MEFF)), all float32 (utils_ros.py:86,105,129)
        ref = arr[..., -1:] + MEFF
        return -np.log((arr + MEFF) / ref)

    # ---- omega (utils_ros.py:78-97) and theta (:99-119): periodic padding
    for name, fmt, unordered in (("omega", "%.5f", True), ("theta", "%.3f", False)):
        arr = npz[name]                                  # float32 (L,L,25)
        nk = arr.shape[2] + 3                            # 28 knots (:81)
        knots = np.linspace(-np.pi - 1.5 * ASTEP, np.pi + 1.5 * ASTEP, nk)
        p_contact = arr[..., 1:].sum(axis=-1)
        e = neg_log_ratio(arr)
        tab = np.concatenate([e[..., 23:25], e[..., 1:25], e[..., 1:3]], axis=-1)
        i, j = np.nonzero(p_contact > PCUT)
        keep = (j > i) if unordered else (j != i)        # (:90) / (:108)
        i, j = i[keep], j[keep]
        xs, x = _fmt_rows(knots, fmt)
        ys, y = _fmt_rows(tab[i, j], fmt)
        out[name] = dict(a=i, b=j, p=p_contact[i, j], x=x, y=y, xs=xs, ys=ys,
                         bin_size=float("%.5f" % ASTEP))

    # ---- phi (utils_ros.py:121-144): mirror padding
    arr = npz["phi"]                                     # float32 (L,L,13)
    nk = arr.shape[2] + 3                                # 16 knots (:124)
    knots = np.linspace(-1.5 * ASTEP, np.pi + 1.5 * ASTEP, nk)
    p_contact = arr[..., 1:].sum(axis=-1)
    e = neg_log_ratio(arr)
    tab = np.concatenate([e[..., 2:0:-1], e[..., 1:13], e[..., 12:10:-1]], axis=-1)
    i, j = np.nonzero(p_contact > PCUT)
    keep = j != i                                        # (:132)
    i, j = i[keep], j[keep]
    xs, x = _fmt_rows(knots, "%.3f")
    ys, y = _fmt_rows(tab[i, j], "%.3f")
    out["phi"] = dict(a=i, b=j, p=p_contact[i, j], x=x, y=y, xs=xs, ys=ys,
                      bin_size=float("%.5f" % ASTEP))
    return out


def text_lines(rec, k):
    """The two lines the reference writes for record k of one type."""
    return ("x_axis\t" + "\t".join(rec["xs"]) + "\n",
            "y_axis\t" + "\t".join(rec["ys"][k]) + "\n")


def select_oracle(rst, sep1, sep2, pcut, seq=None, nogly=False):
    """Restates add_rst's list filters (utils_ros.py:713-723).  Returns dict
    type -> boolean mask over the records of gen_rst_oracle."""
    thr = {"dist": pcut, "omega": pcut + 0.5, "theta": pcut + 0.5, "phi": pcut + 0.6}
    sel = {}
    for name, rec in rst.items():
        a, b, p = rec["a"], rec["b"], rec["p"]
        m = np.zeros(len(a), dtype=bool)
        for k in range(len(a)):  # python-level compare, as the reference does
            ok = abs(int(a[k]) - int(b[k])) >= sep1 and abs(int(a[k]) - int(b[k])) < sep2 and p[k] >= thr[name]
            if nogly and ok:
                ok = seq[a[k]] != "G" and seq[b[k]] != "G"
            m[k] = ok
        sel[name] = m
    return sel
