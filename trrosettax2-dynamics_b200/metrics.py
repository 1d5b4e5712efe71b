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
