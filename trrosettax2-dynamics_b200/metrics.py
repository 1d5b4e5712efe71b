"""Superposition metrics for decoy-quality checks: Kabsch RMSD and a
TM-score restated from the published algorithm (Zhang & Skolnick 2004: d0 = 1.24*(L-15)^(1/3)-1.8,
fragment seeds + iterative re-superposition on residues closer than a growing cutoff).
The reference shells out to the prebuilt bin/TMscore (utils_trX2dy/evaluate_utils.py:58-62);
this is the in-process equivalent for CA traces with identical residue numbering."""
from __future__ import annotations

import numpy as np


def kabsch(P, Q):
    """Rotation R and translation t minimising |R P + t - Q| (rows are points)."""
    pc, qc = P.mean(0), Q.mean(0)
    H = (P - pc).T @ (Q - qc)
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    D = np.diag([1.0, 1.0, d])
    R = Vt.T @ D @ U.T
    return R, qc - R @ pc


def rmsd(P, Q):
    R, t = kabsch(P, Q)
    return float(np.sqrt(np.mean(np.sum((P @ R.T + t - Q) ** 2, axis=1))))


def tm_score(model, native, lnorm=None):
    """TM-score of `model` against `native` (both (L,3) CA traces, same residues),
    normalised by len(native) unless lnorm is given."""
    L = len(native)
    ln = lnorm or L
    d0 = max(0.5, 1.24 * (ln - 15) ** (1.0 / 3.0) - 1.8) if ln > 21 else 0.5
    best = 0.0

    def score(R, t):
        d2 = np.sum((model @ R.T + t - native) ** 2, axis=1)
        return float(np.sum(1.0 / (1.0 + d2 / d0 ** 2)) / ln), d2

    frag = L
    while frag >= 4:
        step = max(1, frag // 2)
        for start in range(0, L - frag + 1, step):
            idx = np.arange(start, start + frag)
            for it in range(20):
                R, t = kabsch(model[idx], native[idx])
                s, d2 = score(R, t)
                best = max(best, s)
                cut = d0 + 1.0 if it == 0 else min(d0 + 1.0 + 0.5 * it, 8.0)
                new = np.nonzero(d2 < cut ** 2)[0]
                while len(new) < 3:
                    cut += 0.5
                    new = np.nonzero(d2 < cut ** 2)[0]
                if len(new) == len(idx) and np.all(new == idx):
                    break
                idx = new
        frag //= 2
    return best


# ---- all pairs of a decoy set on the GPU (libtrx2dyn.so: csrc/metrics.cu) ---------------------

def virtual_cb(n, ca, c):
    """CB rebuilt from N, CA, C (utils_trX2dy/utils.py:132-135); arrays (..., 3)."""
    b, cc = ca - n, c - ca
    return -0.58273431 * np.cross(b, cc) + 0.56802827 * b - 0.54067466 * cc + ca


def glocon_matrix(ctx, cb, dmax=20.0, thr=3.0):
    """get_glocon_matrix (utils_trX2dy/utils.py:543-567) for M decoys at once: cb (M, L, 3) -> (M, M)."""
    import ctypes as C
    from . import capi
    cb = np.ascontiguousarray(cb, dtype=np.float64)
    M, L = cb.shape[:2]
    out = np.zeros((M, M))
    capi.check(capi.lib().trx_glocon_matrix(ctx._h, C.c_int(M), C.c_int(L), cb.ctypes.data_as(C.POINTER(C.c_double)),
                                            C.c_double(dmax), C.c_double(thr), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def tmscore_matrix(ctx, ca):
    """get_tmscore_and_rmsd_matrix (utils_trX2dy/utils.py:514-540) without the per-pair subprocess:
    ca (M, L, 3) -> (tm (M, M), rmsd (M, M)); tm[i, j] is normalised by L (all decoys share the sequence)."""
    import ctypes as C
    from . import capi
    ca = np.ascontiguousarray(ca, dtype=np.float64)
    M, L = ca.shape[:2]
    tm, rm = np.zeros((M, M)), np.zeros((M, M))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    capi.check(capi.lib().trx_tmscore_matrix(ctx._h, C.c_int(M), C.c_int(L), P(ca), P(tm), P(rm)))
    return tm, rm


def kmeans_clusters(matrix, names, n_clusters=10):
    """kmeans_clustering (utils_trX2dy/utils.py:570-580): KMeans(n_clusters, n_init=10, random_state=0) on the
    rows of a GloCon / TM / RMSD matrix; returns {label: [names]} in first-seen order like the reference."""
    from sklearn.cluster import KMeans
    labels = KMeans(n_clusters=n_clusters, n_init=10, random_state=0).fit(matrix).labels_
    clusters = {}
    for i, label in enumerate(labels):
        clusters.setdefault(int(label), []).append(names[i])
    return clusters
