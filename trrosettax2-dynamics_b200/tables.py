"""Restraint-table construction on the host (product path).

Mirrors the reference's interface for this step:
  gen_rst(npz, params)            <- folding/utils_ros/utils_ros.py:6-146
  select(rst, sep1, sep2, params) <- folding/utils_ros/utils_ros.py:706-723 (add_rst filters)
  spline_knots(...)               <- what Rosetta's SplineFunc builds from one text file
                                     (end-knot rule H1 default, H2 optional)

Differences from the reference are only mechanical: no text files are written to
/dev/shm (the knots are rounded to the same decimals the text would carry and kept
as float64 arrays), and records are arrays instead of python lists.  The dtype flow
(float32 probabilities, float32 angle energies, float64 distance energies) is the
reference's, because it decides the printed decimals.
"""
from __future__ import annotations

import json
import os

import numpy as np

TYPES = ("dist", "omega", "theta", "phi")
_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "folding", "data")


def load_params(path=None):
    """folding/data/params.json (same keys as the reference's)."""
    with open(path or os.path.join(_DATA, "params.json")) as fh:
        return json.load(fh)


def round_decimals(v, nd):
    """float(('%%.%df' %% nd) %% v) for every element, vectorised and exact.

    rint(v*10^nd)/10^nd equals the correctly rounded decimal except when v*10^nd
    sits within rounding error of a tie; those few go through string formatting."""
    v = np.asarray(v, dtype=np.float64)
    s = 10.0 ** nd
    t = v * s
    q = np.rint(t)
    near_tie = np.abs(np.abs(t - q) - 0.5) < 1e-6
    out = q / s
    if np.any(near_tie):
        fmt = "%%.%df" % nd
        idx = np.nonzero(near_tie)
        out[idx] = [float(fmt % x) for x in v[idx]]
    return out


def gen_rst(npz, params, use_orient=None):
    """Distograms -> restraint records.  Returns {type: {a, b, p, x, y, bin_size}}
    with a,b int32 0-based (reference order), p float32, x (K,) and y (n,K) float64
    knots as Rosetta would parse them from the reference's text files."""
    if use_orient is None:
        use_orient = params.get("USE_ORIENT", True) in (True, "True")
    meff, dcut, alpha = params["MEFF"], params["DCUT"], params["ALPHA"]
    astep = np.deg2rad(params["ASTEP"])
    pcut = 0.05  # the reference hard-codes this here (utils_ros.py:18); -pd only acts in select()
    rst = {}

    d = npz["dist"]
    centres = 4.25 + params["DSTEP"] * np.arange(32)
    p = d[..., 5:].sum(axis=-1)
    e = params["EBASE"] - np.log((d[..., 5:] + meff) / (d[..., 36:37] * (centres / dcut) ** alpha + 1e-6))
    rep = np.maximum(e[..., 0], 0.0)[..., None] + np.asarray(params["EREP"])
    i, j = np.nonzero(p > pcut)
    m = j > i
    i, j = i[m], j[m]
    rst["dist"] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j],
                       x=round_decimals(np.concatenate([params["DREP"], centres]), 3),
                       y=round_decimals(np.concatenate([rep[i, j], e[i, j]], axis=-1), 3),
                       bin_size=0.5)
    if not use_orient:
        return rst

    astep5 = float(round_decimals(astep, 5))
    for name, nd, unordered in (("omega", 5, True), ("theta", 3, False), ("phi", 3, False)):
        t = npz[name]
        nb = t.shape[2]
        p = t[..., 1:].sum(axis=-1)
        e = -np.log((t + meff) / (t[..., -1:] + meff))          # float32 throughout
        if name == "phi":
            lo = -1.5 * astep
            e = np.concatenate([e[..., 2:0:-1], e[..., 1:], e[..., nb - 1:nb - 3:-1]], axis=-1)
        else:
            lo = -np.pi - 1.5 * astep
            e = np.concatenate([e[..., nb - 2:], e[..., 1:], e[..., 1:3]], axis=-1)
        i, j = np.nonzero(p > pcut)
        m = (j > i) if unordered else (j != i)
        i, j = i[m], j[m]
        rst[name] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j],
                         x=round_decimals(np.linspace(lo, np.pi + 1.5 * astep, nb + 3), nd),
                         y=round_decimals(e[i, j], nd), bin_size=astep5)
    return rst


def select(rst, sep1, sep2, params, seq=None, nogly=False):
    """add_rst's filters: sequence separation in [sep1, sep2) and probability
    >= PCUT (dist), PCUT+0.5 (omega, theta), PCUT+0.6 (phi).  Returns masks."""
    pcut = params["PCUT"]
    thr = {"dist": pcut, "omega": pcut + 0.5, "theta": pcut + 0.5, "phi": pcut + 0.6}
    out = {}
    for name, rec in rst.items():
        sep = np.abs(rec["a"] - rec["b"])
        m = (sep >= sep1) & (sep < sep2) & (rec["p"] >= np.float32(thr[name]))
        if nogly:
            g = np.frombuffer(seq.encode(), dtype=np.uint8) == ord("G")
            m &= ~g[rec["a"]] & ~g[rec["b"]]
        out[name] = m
    return out


def spline_knots(x, y, bin_size, rule="H1"):
    """Knots of Rosetta's SplineFunc for listed points (x, y[n,K]).
    H1: the spline's bounds are x_1-bin_size / x_n+bin_size carrying y_1 / y_n,
    all listed points interior (K+2 knots).  H2: bounds on the end points (K knots)."""
    if rule == "H2":
        return np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    if rule != "H1":
        raise ValueError("end-knot rule must be 'H1' or 'H2'")
    xx = np.concatenate([[x[0] - bin_size], x, [x[-1] + bin_size]])
    yy = np.concatenate([y[:, :1], y, y[:, -1:]], axis=1)
    return np.ascontiguousarray(xx), np.ascontiguousarray(yy)


def active_restraints(rst, masks=None, rule="H1"):
    """Flat arrays for the C-ABI: {type: (a, b, x, y)} of the selected records."""
    out = {}
    for name in TYPES:
        if name not in rst:
            continue
        rec = rst[name]
        m = slice(None) if masks is None else masks[name]
        x, y = spline_knots(rec["x"], rec["y"][m], rec["bin_size"], rule)
        out[name] = (np.ascontiguousarray(rec["a"][m], dtype=np.int32),
                     np.ascontiguousarray(rec["b"][m], dtype=np.int32), x, y)
    return out
