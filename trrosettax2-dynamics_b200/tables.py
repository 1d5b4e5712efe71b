"""Restraint-table construction on the host (product path).

Mirrors the reference's interface for this step:
  gen_rst(npz, params)            <- folding/utils_ros/utils_ros.py:6-146, and its -r variants
                                     gen_idp_rst (:196-373), gen_rst_af2 (:148-194), gen_gpcr_rst (:484-654)
  select(rst, sep1, sep2, params) <- folding/utils_ros/utils_ros.py:706-723 (add_rst filters)
  select_idr(rst, idr, params)    <- folding/utils_ros/utils_ros.py:745-760 (add_idr_rst filters, mode 3)
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
        return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}


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


def _dist_table(d, params, centres):
    """35-knot distance table of every pair (utils_ros.py:56-60): float32 numerator, float64 denominator."""
    e = params["EBASE"] - np.log((d[..., 5:] + params["MEFF"]) / (d[..., -1:] * (centres / params["DCUT"]) ** params["ALPHA"] + 1e-6))
    rep = np.maximum(e[..., 0], 0.0)[..., None] + np.asarray(params["EREP"])
    return np.concatenate([rep, e], axis=-1)


def _pad(e, name):
    """Periodic padding [E23,E24;E1..E24;E1,E2] of omega/theta, mirror padding [E2,E1;E1..E12;E12,E11] of phi."""
    nb = e.shape[-1]
    if name == "phi":
        return np.concatenate([e[..., 2:0:-1], e[..., 1:], e[..., nb - 1:nb - 3:-1]], axis=-1)
    return np.concatenate([e[..., nb - 2:], e[..., 1:], e[..., 1:3]], axis=-1)


def _template_histogram(values, edges, nbins, gate):
    """pros + get_sample of the gpcr variant (utils_ros.py:395-482): bin M template maps, then spread
    every vote as a Gaussian over bin indices (std 0.5 / 1.0 / 1.5 by how many templates agree)."""
    J = (edges[None, None, None, :] < values[..., None]).sum(-1)
    J = np.where(J >= nbins, 0, J) if gate is None else np.where(gate == 0, 0, J)
    M = values.shape[0]
    count = np.eye(nbins)[J].sum(axis=0)                                  # (L, L, nbins) votes per bin
    std = np.where(count < M / 3, 1.5, np.where(count > 2 * M / 3, 0.5, 1.0))
    x = np.arange(nbins)
    out = np.zeros(count.shape)
    for k in range(nbins):                                                # same accumulation order as the reference (k ascending)
        g = 1 / np.sqrt(2 * np.pi * std[..., k] ** 2)
        pdf = g[..., None] * np.exp(-((x[None, None, :] - k) ** 2) / (2 * std[..., k, None] ** 2))
        for _ in range(int(count[..., k].max())):
            out += np.where((count[..., k] > _)[..., None], pdf, 0.0)
    return out / M, J


def _blend(tab, tmpl, knots, mask, rg=5):
    """ling_sumlt (utils_ros.py:375-394) on the masked pairs: the rg lowest-energy knots of the template
    table are replaced in `tab` by the straight line between the knots just outside them."""
    t = tab.copy()
    ii, jj = np.nonzero(mask)
    if len(ii) == 0:
        return t
    idx = np.stack([np.argsort(tmpl[i, j])[:rg] for i, j in zip(ii, jj)])          # the reference's (unstable) argsort, row by row
    lo, hi = idx.min(axis=1) - 1, idx.max(axis=1) + 1
    lo = np.where(lo < 0, lo + 1, lo)
    hi = np.where(hi >= len(knots), hi - 1, hi)
    rows = t[ii, jj]
    n = np.arange(len(ii))
    tl, th = rows[n, lo], rows[n, hi]
    val = (knots[idx] - knots[hi][:, None]) / (knots[lo] - knots[hi])[:, None] * (tl - th)[:, None] + th[:, None]
    rows[n[:, None], idx] = val                                                    # cast back to the table's dtype (float32 for angles)
    t[ii, jj] = rows
    return t


def gen_rst(npz, params, use_orient=None, variant="no-idp", known=None):
    """Distograms -> restraint records.  Returns {type: {a, b, p, x, y, bin_size}}
    with a,b int32 0-based (reference order), p float32, x (K,) and y (n,K) float64
    knots as Rosetta would parse them from the reference's text files.

    variant (folding.py -r):  'no-idp' gen_rst (utils_ros.py:6-146) | 'idp' gen_idp_rst (:196-373, pairs
    flagged in npz['idr'] take the most probable bin as energy reference) | 'gpcr' gen_gpcr_rst (:484-654,
    tables blended with a histogram of the template maps in `known` on the flagged pairs) |
    'af2' gen_rst_af2 (:148-194, 64-bin CA-CA distogram, distance only)."""
    if use_orient is None:
        use_orient = params.get("USE_ORIENT", True) in (True, "True")
    if variant == "af2":
        if use_orient:
            raise RuntimeError("AF2 Not support ")          # the reference's own error (utils_ros.py:150)
        return _gen_rst_af2(npz, params)
    if variant not in ("no-idp", "idp", "gpcr"):
        raise ValueError("unknown restraint variant %r" % (variant,))
    if variant == "gpcr" and known is None:
        raise ValueError("the gpcr variant needs the template maps (-KNOWN)")
    meff = params["MEFF"]
    astep = np.deg2rad(params["ASTEP"])
    pcut = 0.05  # the reference hard-codes this here (utils_ros.py:18); -pd only acts in select()
    if variant == "no-idp":
        return _gen_rst_plain(npz, params, use_orient, meff, astep, pcut)
    idr = np.asarray(npz["idr"]).astype(bool)
    rst = {}

    d = npz["dist"]
    centres = 4.25 + params["DSTEP"] * np.arange(32)
    knots = np.concatenate([params["DREP"], centres])
    p = d[..., 5:].sum(axis=-1)
    tab = _dist_table(d, params, centres)
    gate = None
    if variant == "idp":
        top = d[..., 5:].max(axis=-1)
        bk = (centres[None, None, :] / centres[d[..., 5:].argmax(axis=-1)][..., None]) ** params["ALPHA"]
        e_i = params["EBASE"] - np.log((d[..., 5:] + meff) / (top[..., None] * bk + 1e-6))
        tab = np.where(idr[..., None], np.concatenate([tab[..., :3], e_i], axis=-1), tab)   # repulsive knots from the last-bin table
    elif variant == "gpcr":
        hist, gate = _template_histogram(np.asarray(known["dist"]), np.arange(2, 20.5, 0.5), 37, None)
        tab = _blend(tab, _dist_table(hist, params, centres), knots, idr)
    i, j = np.nonzero(p > pcut)
    m = j > i
    i, j = i[m], j[m]
    rst["dist"] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j], x=round_decimals(knots, 3),
                       y=round_decimals(tab[i, j], 3), bin_size=0.5)
    if not use_orient:
        return rst

    astep5 = float(round_decimals(astep, 5))
    for name, nd, unordered in (("omega", 5, True), ("theta", 3, False), ("phi", 3, False)):
        t = npz[name]
        nb = t.shape[2]
        lo = -1.5 * astep if name == "phi" else -np.pi - 1.5 * astep
        knots = np.linspace(lo, np.pi + 1.5 * astep, nb + 3)
        p = t[..., 1:].sum(axis=-1)
        e = _pad(-np.log((t + meff) / (t[..., -1:] + meff)), name)          # float32 throughout
        if variant == "idp":
            e_i = _pad(-np.log((t + meff) / (t.max(axis=-1) + meff)[..., None]), name)
            e = np.where(idr[..., None], e_i, e)
        elif variant == "gpcr":
            # templates: omega, theta_asym; phi is binned from the THETA values (utils_ros.py:433, kept)
            src = np.asarray(known["omega" if name == "omega" else "theta_asym"])
            edges = np.arange(0, np.pi, np.pi / 12) if name == "phi" else np.arange(-np.pi, np.pi, np.pi / 12)
            hist, _ = _template_histogram(src, edges, nb, gate)
            e = _blend(e, _pad(-np.log((hist + meff) / (hist[..., -1:] + meff)), name), knots, idr)
        i, j = np.nonzero(p > pcut)
        m = (j > i) if unordered else (j != i)
        i, j = i[m], j[m]
        rst[name] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j],
                         x=round_decimals(knots, nd), y=round_decimals(e[i, j], nd), bin_size=astep5)
    return rst


def _gen_rst_plain(npz, params, use_orient, meff, astep, pcut):
    """gen_rst proper (utils_ros.py:6-146), the variant the dynamics loop calls once per iteration: the pairs are
    selected first and the energies evaluated for those rows only (elementwise the same float32 / float64
    operations as on the full (L, L, bins) arrays: identical knots, ~5 x less work at L = 300)."""
    rst = {}
    d = npz["dist"]
    centres = 4.25 + params["DSTEP"] * np.arange(32)
    p = d[..., 5:].sum(axis=-1)
    i, j = np.nonzero(p > pcut)
    m = j > i
    i, j = i[m], j[m]
    rst["dist"] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j],
                       x=round_decimals(np.concatenate([params["DREP"], centres]), 3),
                       y=round_decimals(_dist_table(d[i, j], params, centres), 3), bin_size=0.5)
    if not use_orient:
        return rst
    astep5 = float(round_decimals(astep, 5))
    for name, nd, unordered in (("omega", 5, True), ("theta", 3, False), ("phi", 3, False)):
        t = npz[name]
        nb = t.shape[2]
        lo = -1.5 * astep if name == "phi" else -np.pi - 1.5 * astep
        p = t[..., 1:].sum(axis=-1)
        i, j = np.nonzero(p > pcut)
        m = (j > i) if unordered else (j != i)
        i, j = i[m], j[m]
        ts = t[i, j]
        e = _pad(-np.log((ts + meff) / (ts[..., -1:] + meff)), name)          # float32 throughout
        rst[name] = dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j],
                         x=round_decimals(np.linspace(lo, np.pi + 1.5 * astep, nb + 3), nd), y=round_decimals(e, nd),
                         bin_size=astep5)
    return rst


def _gen_rst_af2(npz, params):
    """gen_rst_af2 (utils_ros.py:148-194).  AtomPair restraints sit on CA (record key 'atom').  Quirk kept:
    every bin uses the LAST bin's background term (bkgr[None,None,-1], :172)."""
    d, edges = npz["dist"], np.asarray(npz["bins"])
    bins = edges[5:-1]
    p = d[..., 6:-1].sum(axis=-1)
    bk_last = ((bins / params["DCUT"]) ** params["ALPHA"])[-1]
    e = params["EBASE"] - np.log((d[..., 6:-1] + params["MEFF"]) / (d[..., -2][..., None] * bk_last + 1e-6))
    rep = np.maximum(e[..., 0], 0.0)[..., None] + np.asarray(params["EREP"])
    tab = np.concatenate([rep, e], axis=-1)
    knots = np.concatenate([[0.0, 2.325, 3.575], bins])
    if tab.shape[-1] != 60:
        raise ValueError("the af2 variant expects a 64-bin distogram and 63 bin edges (60 knots), got %d knots" % tab.shape[-1])
    i, j = np.nonzero(p > 0.0025)
    m = j > i
    i, j = i[m], j[m]
    return {"dist": dict(a=i.astype(np.int32), b=j.astype(np.int32), p=p[i, j], x=round_decimals(knots, 3),
                         y=round_decimals(tab[i, j], 3), bin_size=0.3125, atom="CA")}


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


def select_idr(rst, idr, params, seq=None, nogly=False):
    """add_idr_rst's filters (utils_ros.py:745-760, mode 3): pairs flagged in `idr` at any sequence
    separation, above the probability thresholds of add_rst.  Returns masks."""
    pcut = params["PCUT"]
    thr = {"dist": pcut, "omega": pcut + 0.5, "theta": pcut + 0.5, "phi": pcut + 0.6}
    flag = np.asarray(idr) != 0
    out = {}
    for name, rec in rst.items():
        m = flag[rec["a"], rec["b"]] & (rec["p"] >= np.float32(thr[name]))
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
    """Flat arrays for the C-ABI: {type: (a, b, x, y)} of the selected records (+ 'dist_atom': 'CA' for
    the af2 variant, whose AtomPair restraints sit on CA)."""
    out = {}
    if rst.get("dist", {}).get("atom") == "CA":
        out["dist_atom"] = "CA"
    for name in TYPES:
        if name not in rst:
            continue
        rec = rst[name]
        m = slice(None) if masks is None else masks[name]
        x, y = spline_knots(rec["x"], rec["y"][m], rec["bin_size"], rule)
        out[name] = (np.ascontiguousarray(rec["a"][m], dtype=np.int32),
                     np.ascontiguousarray(rec["b"][m], dtype=np.int32), x, y)
    return out
