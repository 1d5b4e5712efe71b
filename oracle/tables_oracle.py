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


# ---------------------------------------------------------------------------------------------
# Restraint variants: -r idp / af2 / gpcr and the mode-3 selection.
# Restated with explicit loops where the reference loops, so that this file stays a plain
# reading of utils_ros.py; the product (trx2dyn.tables) is vectorised and tested against it.

def _pack(i, j, p, knots, tab, fmt, bin_size):
    xs, x = _fmt_rows(knots, fmt)
    ys, y = _fmt_rows(tab, fmt)
    return dict(a=i, b=j, p=p, x=x, y=y, xs=xs, ys=ys, bin_size=float("%.5f" % bin_size))


def _periodic_pad(e):      # [E23,E24 ; E1..E24 ; E1,E2]  (utils_ros.py:87)
    return np.concatenate([e[..., -2:], e[..., 1:], e[..., 1:3]], axis=-1)


def _mirror_pad(e):        # [E2,E1 ; E1..E12 ; E12,E11]  (utils_ros.py:130)
    return np.concatenate([np.flip(e[..., 1:3], axis=-1), e[..., 1:], np.flip(e[..., -2:], axis=-1)], axis=-1)


def gen_idp_rst_oracle(npz, use_orient=True, params=PARAMS):
    """gen_idp_rst (utils_ros.py:196-373): as gen_rst, but pairs flagged in npz['idr'] take their
    energies relative to the MOST PROBABLE bin instead of the last one."""
    MEFF, DCUT, ALPHA, EBASE = params["MEFF"], params["DCUT"], params["ALPHA"], params["EBASE"]
    EREP, DREP, DSTEP = params["EREP"], params["DREP"], params["DSTEP"]
    ASTEP = np.deg2rad(params["ASTEP"])
    PCUT = 0.05
    idr = np.asarray(npz["idr"])
    out = {}
    dist = npz["dist"]
    bins = np.array([4.25 + DSTEP * i for i in range(32)])
    prob = np.sum(dist[:, :, 5:], axis=-1)
    top = np.argmax(dist[:, :, 5:], axis=-1)
    idr_bkgr = (bins[None, None, :] / bins[top][:, :, None]) ** ALPHA                     # (:249)
    idr_attr = -np.log((dist[:, :, 5:] + MEFF) / (np.max(dist[:, :, 5:], axis=-1)[:, :, None] * idr_bkgr + 1e-6)) + EBASE
    bkgr = (bins / DCUT) ** ALPHA
    attr = -np.log((dist[:, :, 5:] + MEFF) / (dist[:, :, -1][:, :, None] * bkgr[None, None, :] + 1e-6)) + EBASE
    repul = np.maximum(attr[:, :, 0], 0.0)[:, :, None] + np.array(EREP)[None, None, :]   # from the LAST-bin table for both (:254)
    tab_n = np.concatenate([repul, attr], axis=-1)
    tab_i = np.concatenate([repul, idr_attr], axis=-1)
    knots = np.concatenate([DREP, bins])
    i, j = np.where(prob > PCUT)
    keep = j > i
    i, j = i[keep], j[keep]
    rows = np.array([tab_i[a, b] if idr[a, b] else tab_n[a, b] for a, b in zip(i, j)]).reshape(len(i), 35)
    out["dist"] = _pack(i, j, prob[i, j], knots, rows, "%.3f", 0.5)
    if not use_orient:
        return out
    for name, fmt, unordered, pad in (("omega", "%.5f", True, _periodic_pad), ("theta", "%.3f", False, _periodic_pad),
                                      ("phi", "%.3f", False, _mirror_pad)):
        arr = npz[name]
        nk = arr.shape[2] + 3
        lo = -1.5 * ASTEP if name == "phi" else -np.pi - 1.5 * ASTEP
        knots = np.linspace(lo, np.pi + 1.5 * ASTEP, nk)
        prob = np.sum(arr[:, :, 1:], axis=-1)
        e_i = pad(-np.log((arr + MEFF) / (np.max(arr, axis=-1) + MEFF)[:, :, None]))     # max over ALL bins, bin 0 included
        e_n = pad(-np.log((arr + MEFF) / (arr[:, :, -1] + MEFF)[:, :, None]))
        i, j = np.where(prob > PCUT)
        keep = (j > i) if unordered else (j != i)
        i, j = i[keep], j[keep]
        rows = np.array([e_i[a, b] if idr[a, b] else e_n[a, b] for a, b in zip(i, j)]).reshape(len(i), nk)
        out[name] = _pack(i, j, prob[i, j], knots, rows, fmt, ASTEP)
    return out


def gen_rst_af2_oracle(npz, params=PARAMS):
    """gen_rst_af2 (utils_ros.py:148-194): AlphaFold-style 64-bin CA-CA distogram, distance only.
    Quirk kept: the background of EVERY bin is the last bin's (bkgr[None,None,-1], :172)."""
    MEFF, DCUT, ALPHA, EBASE, EREP = params["MEFF"], params["DCUT"], params["ALPHA"], params["EBASE"], params["EREP"]
    PCUT = 0.0025
    dist, af_bins = npz["dist"], np.asarray(npz["bins"])
    bins = af_bins[5:-1]
    prob = np.sum(dist[:, :, 6:-1], axis=-1)
    bkgr = (bins / DCUT) ** ALPHA
    attr = -np.log((dist[:, :, 6:-1] + MEFF) / (dist[:, :, -2][:, :, None] * bkgr[None, None, -1] + 1e-6)) + EBASE
    repul = np.maximum(attr[:, :, 0], 0.0)[:, :, None] + np.array(EREP)[None, None, :]
    tab = np.concatenate([repul, attr], axis=-1)
    knots = np.concatenate([[0.0, 2.325, 3.575], bins])
    assert tab.shape[-1] == 60 and len(knots) == 60     # the reference formats exactly 60 values (:181-187)
    i, j = np.where(prob > PCUT)
    keep = j > i
    i, j = i[keep], j[keep]
    rec = _pack(i, j, prob[i, j], knots, tab[i, j], "%.3f", 0.3125)
    rec["atom"] = "CA"
    return {"dist": rec}


def _bin_templates(known, use_orient):
    """pros (utils_ros.py:395-450): real-valued template maps -> one-hot bins.  Quirk kept: phi is
    binned from the THETA values (:433)."""
    d = np.asarray(known["dist"])
    edges = np.arange(2, 20.5, 0.5)
    J = (edges[None, None, None, :] < d[..., None]).sum(-1)
    J = np.where(J >= 37, 0, J)
    oh = {"dist": np.eye(37)[J]}
    if use_orient:
        ae = np.arange(-np.pi, np.pi, np.pi / 12)
        for name, key in (("omega", "omega"), ("theta", "theta_asym")):
            Ja = (ae[None, None, None, :] < np.asarray(known[key])[..., None]).sum(-1)
            oh[name] = np.eye(25)[np.where(J == 0, 0, Ja)]
        pe = np.arange(0, np.pi, np.pi / 12)
        Jp = (pe[None, None, None, :] < np.asarray(known["theta_asym"])[..., None]).sum(-1)
        oh["phi"] = np.eye(13)[np.where(J == 0, 0, Jp)]
    return oh


def _template_histogram(onehot):
    """get_sample (utils_ros.py:456-482): every template vote is spread as a Gaussian over bin
    indices, narrower (std .5) when > 2/3 of the templates agree, wider (1.5) when < 1/3 do."""
    M, H, W, Cn = onehot.shape
    count = onehot.sum(axis=0)
    out = np.zeros((H, W, Cn))
    x = np.arange(Cn)
    for i in range(H):
        for j in range(W):
            for k in np.where(count[i, j] != 0)[0]:
                c = count[i, j, k]
                std = 1.5 if c < M / 3 else (0.5 if c > 2 * M / 3 else 1.0)
                g = (1 / (np.sqrt(2 * np.pi * std ** 2)) * np.exp(-((x - k) ** 2) / (2 * std ** 2)))
                for _ in range(int(c)):
                    out[i, j, :] += g
    return out / M


def _blend(test, tmpl, knots, mask, rg=5):
    """ling_sumlt (utils_ros.py:375-394): on masked pairs the rg lowest-energy knots of the TEMPLATE
    table are replaced, in the predicted table, by the straight line between the knots just outside
    them.  Ties: the padded angular tables repeat knots, so ties are common; the reference calls
    numpy's default (unstable) argsort, whose tie order depends on the numpy build / CPU -- the same
    call is made here, and tests accept either order on rows whose 5th/6th lowest values tie."""
    t = test.copy()
    for i in range(test.shape[0]):
        for j in range(test.shape[1]):
            if mask[i, j]:
                idx = np.argsort(tmpl[i, j])[:rg]
                lo, hi = idx.min() - 1, idx.max() + 1
                if lo < 0:
                    lo += 1
                if hi >= len(knots):
                    hi -= 1
                t[i, j][idx] = (knots[idx] - knots[hi]) / (knots[lo] - knots[hi]) * (t[i, j][lo] - t[i, j][hi]) + t[i, j][hi]
    return t


def gen_gpcr_rst_oracle(npz, known, use_orient=True, params=PARAMS):
    """gen_gpcr_rst (utils_ros.py:484-654): predicted tables blended with a histogram of known
    (template) structures on the pairs flagged in npz['idr']."""
    MEFF, DCUT, ALPHA, EBASE = params["MEFF"], params["DCUT"], params["ALPHA"], params["EBASE"]
    EREP, DREP, DSTEP = params["EREP"], params["DREP"], params["DSTEP"]
    ASTEP = np.deg2rad(params["ASTEP"])
    PCUT = 0.05
    idr = np.asarray(npz["idr"])
    oh = _bin_templates(known, use_orient)
    out = {}
    dist = npz["dist"]
    bins = np.array([4.25 + DSTEP * i for i in range(32)])
    prob = np.sum(dist[:, :, 5:], axis=-1)
    bkgr = (bins / DCUT) ** ALPHA

    def table(d):
        attr = -np.log((d[:, :, 5:] + MEFF) / (d[:, :, -1][:, :, None] * bkgr[None, None, :] + 1e-6)) + EBASE
        repul = np.maximum(attr[:, :, 0], 0.0)[:, :, None] + np.array(EREP)[None, None, :]
        return np.concatenate([repul, attr], axis=-1)
    knots = np.concatenate([DREP, bins])
    tab = _blend(table(dist), table(_template_histogram(oh["dist"])), knots, idr)
    i, j = np.where(prob > PCUT)
    keep = j > i
    i, j = i[keep], j[keep]
    out["dist"] = _pack(i, j, prob[i, j], knots, tab[i, j], "%.3f", 0.5)
    if not use_orient:
        return out
    for name, fmt, unordered, pad in (("omega", "%.5f", True, _periodic_pad), ("theta", "%.3f", False, _periodic_pad),
                                      ("phi", "%.3f", False, _mirror_pad)):
        arr = npz[name]
        nk = arr.shape[2] + 3
        lo = -1.5 * ASTEP if name == "phi" else -np.pi - 1.5 * ASTEP
        knots = np.linspace(lo, np.pi + 1.5 * ASTEP, nk)
        prob = np.sum(arr[:, :, 1:], axis=-1)
        e = pad(-np.log((arr + MEFF) / (arr[:, :, -1] + MEFF)[:, :, None]))               # float32
        cate = _template_histogram(oh[name])
        ce = pad(-np.log((cate + MEFF) / (cate[:, :, -1] + MEFF)[:, :, None]))            # float64
        tab = _blend(e, ce, knots, idr)                                                    # stays float32 (t = test.copy())
        i, j = np.where(prob > PCUT)
        keep = (j > i) if unordered else (j != i)
        i, j = i[keep], j[keep]
        out[name] = _pack(i, j, prob[i, j], knots, tab[i, j], fmt, ASTEP)
    return out


def select_idr_oracle(rst, idr, pcut, seq=None, nogly=False):
    """add_idr_rst's list filters (utils_ros.py:745-760): pairs flagged in `idr` (any sequence
    separation) above the probability thresholds of add_rst."""
    thr = {"dist": pcut, "omega": pcut + 0.5, "theta": pcut + 0.5, "phi": pcut + 0.6}
    sel = {}
    for name, rec in rst.items():
        m = np.zeros(len(rec["a"]), dtype=bool)
        for k, (a, b, p) in enumerate(zip(rec["a"], rec["b"], rec["p"])):
            ok = bool(idr[a, b]) and p >= thr[name]
            if nogly and ok:
                ok = seq[a] != "G" and seq[b] != "G"
            m[k] = ok
        sel[name] = m
    return sel
