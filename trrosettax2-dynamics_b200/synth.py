"""Synthetic targets for tests and bench: a seeded compact backbone and
the distance / orientation distograms a network would predict for it, in the npz
format the reference's folding.py reads (dist 37 bins, omega/theta 25, phi 13; bin 0 =
'no contact / beyond 20 A'; utils_trX2dy/utils.py:185-249 defines the binning)."""
from __future__ import annotations

import numpy as np


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def ca_trace(L, rng, n_domains=1):
    """Self-avoiding chain growth (3.8 A steps, >= 4.2 A between non-neighbours) confined
    to one sphere per domain (radius ~ 2.2*L^0.38*1.3 A); domains joined by a short linker."""
    per = [L // n_domains + (1 if d < L % n_domains else 0) for d in range(n_domains)]
    pts = []
    centre = np.zeros(3)
    for d, n in enumerate(per):
        radius = 2.2 * n ** 0.38 * 1.3
        start = centre + _unit(rng.normal(size=3)) * radius * 0.5 if not pts else pts[-1] + _unit(centre - pts[-1]) * 3.8
        dom = [start]
        direction = _unit(rng.normal(size=3))
        while len(dom) < n:
            placed = False
            for attempt in range(200):
                # persistent direction with noise gives helix/strand-like runs
                cand_dir = _unit(direction + rng.normal(size=3) * (0.6 if attempt < 100 else 1.5))
                cand = dom[-1] + 3.8 * cand_dir
                if np.linalg.norm(cand - centre) > radius:
                    continue
                others = np.array(pts + dom[:-1]) if (pts or len(dom) > 1) else np.zeros((0, 3))
                if len(others) and np.min(np.linalg.norm(others - cand, axis=1)) < 4.2:
                    continue
                dom.append(cand)
                direction = cand_dir
                placed = True
                break
            if not placed:  # dead end: back up a few residues
                del dom[max(1, len(dom) - 5):]
                direction = _unit(rng.normal(size=3))
        pts += dom
        centre = centre + _unit(rng.normal(size=3)) * radius * 2.3
    return np.array(pts[:L])


def backbone_from_ca(ca):
    """N, CA, C, CB placed from the CA trace (approximate ideal geometry; CB by the
    virtual-CB formula the reference uses for Gly, utils_trX2dy/utils.py:132-135)."""
    L = len(ca)
    prev = np.vstack([2 * ca[0] - ca[1], ca[:-1]])
    nxt = np.vstack([ca[1:], 2 * ca[-1] - ca[-2]])
    b, c = _unit(ca - prev), _unit(nxt - ca)
    nrm = np.cross(b, c)
    bad = np.linalg.norm(nrm, axis=1) < 1e-3
    nrm[bad] = np.cross(b[bad], np.array([0.3, 0.5, 0.8]))
    nrm = _unit(nrm)
    out = _unit(b - c + 1e-3 * nrm)
    n_at = ca + 1.458 * _unit(-0.85 * b + 0.35 * out + 0.25 * nrm)
    c_at = ca + 1.523 * _unit(0.85 * c + 0.35 * out - 0.25 * nrm)
    bb, cc = ca - n_at, c_at - ca
    cb = -0.58273431 * np.cross(bb, cc) + 0.56802827 * bb - 0.54067466 * cc + ca
    xyz = np.stack([n_at, ca, c_at, cb], axis=1)
    assert xyz.shape == (L, 4, 3)
    return xyz


def _dihedral(p1, p2, p3, p4):
    b0, b1, b2 = p1 - p2, _unit(p3 - p2), p4 - p3
    v = b0 - np.sum(b0 * b1, -1, keepdims=True) * b1
    w = b2 - np.sum(b2 * b1, -1, keepdims=True) * b1
    return np.arctan2(np.sum(np.cross(b1, v) * w, -1), np.sum(v * w, -1))


def six_d(xyz):
    """d, omega, theta, phi (L,L) from N,CA,C,CB coordinates (trRosetta definitions)."""
    n, ca, cb = xyz[:, 0], xyz[:, 1], xyz[:, 3]
    L = len(ca)
    i, j = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    d = np.linalg.norm(cb[i] - cb[j], axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        omega = _dihedral(ca[i], cb[i], cb[j], ca[j])
        theta = _dihedral(n[i], ca[i], cb[i], cb[j])
        u, v = _unit(ca[i] - cb[i]), _unit(cb[j] - cb[i])
        phi = np.arccos(np.clip(np.sum(u * v, -1), -1, 1))
    for a in (omega, theta, phi):
        a[np.arange(L), np.arange(L)] = 0.0
    return d, omega, theta, phi


def _gauss_bins(centre_bin, nb, sigma, rng_uniform_mass=0.02, bin0=0.02):
    """rows of probabilities over bins 0..nb-1: gaussian over bins 1.. around centre_bin."""
    k = np.arange(1, nb)[None, None, :]
    g = np.exp(-0.5 * ((k - centre_bin[..., None]) / sigma[..., None]) ** 2)
    g /= g.sum(-1, keepdims=True)
    g = (1 - rng_uniform_mass) * g + rng_uniform_mass / (nb - 1)
    return np.concatenate([np.full(g.shape[:2] + (1,), bin0), (1 - bin0) * g], axis=-1)


def distograms(xyz, rng, dense=False, periodic_sigma=True):
    """npz-like dict (float32) for the backbone.  dense=True makes every pair a contact
    (distances clipped below 19.4 A) so that all L(L-1)/2 pairs carry all restraints."""
    d, omega, theta, phi = six_d(xyz)
    L = len(d)
    if dense:
        d = np.minimum(d, 19.4)
    contact = d < 20.0
    np.fill_diagonal(contact, False)
    sig = rng.uniform(0.7, 2.5, size=(L, L))
    sig = np.triu(sig, 1) + np.triu(sig, 1).T + np.eye(L)

    def far(nb):
        row = np.full(nb, 0.05 / (nb - 1))
        row[0] = 0.95
        return row

    out = {}
    dbin = np.clip((d - 2.0) / 0.5, 0, 35.999) + 1.0   # continuous bin coordinate, centre = k+0.5 -> use centres
    p = _gauss_bins(np.floor(dbin) + 0.0, 37, sig)
    p = np.where(contact[..., None], p, far(37))
    out["dist"] = 0.5 * (p + p.transpose(1, 0, 2))
    for name, val, nb, lo, sym in (("omega", omega, 25, -np.pi, True), ("theta", theta, 25, -np.pi, False),
                                   ("phi", phi, 13, 0.0, False)):
        step = np.deg2rad(15.0)
        kb = np.clip(np.floor((val - lo) / step), 0, nb - 2) + 1.0
        k = np.arange(1, nb)[None, None, :]
        diff = k - kb[..., None]
        if name != "phi":  # periodic distance in bin space
            diff = (diff + (nb - 1) / 2) % (nb - 1) - (nb - 1) / 2
        g = np.exp(-0.5 * (diff / sig[..., None]) ** 2)
        g /= g.sum(-1, keepdims=True)
        g = 0.98 * g + 0.02 / (nb - 1)
        p = np.concatenate([np.full((L, L, 1), 0.02), 0.98 * g], axis=-1)
        p = np.where(contact[..., None], p, far(nb))
        if sym:
            p = 0.5 * (p + p.transpose(1, 0, 2))
        out[name] = p
    for key in out:
        a = out[key]
        out[key] = (a / a.sum(-1, keepdims=True)).astype(np.float32)
    return out


def hinge(xyz, rng, angle_deg=25.0):
    """Second conformation: the C-terminal half rotated about an axis through the hinge."""
    L = len(xyz)
    h = L // 2
    axis = _unit(rng.normal(size=3))
    t = np.deg2rad(angle_deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * K @ K
    out = xyz.copy()
    pivot = xyz[h, 1]
    out[h:] = (xyz[h:] - pivot) @ R.T + pivot
    return out


_B = dict(n_ca=1.458, ca_c=1.523, c_n=1.329)
_A = dict(n_ca_c=np.deg2rad(111.2), ca_c_n=np.deg2rad(116.2), c_n_ca=np.deg2rad(121.7))


def _place(a, b, c, bond, ang, tor):
    """NeRF, vectorised over leading axes: |cd|=bond, angle(b,c,d)=ang, dihedral(a,b,c,d)=tor."""
    bc = _unit(c - b)
    n = _unit(np.cross(b - a, bc))
    m = np.cross(n, bc)
    tor = np.asarray(tor)[..., None]
    return c + (-bond * np.cos(ang)) * bc + (bond * np.sin(ang) * np.cos(tor)) * m + (bond * np.sin(ang) * np.sin(tor)) * n


def _extend(N, CA, C, tors_prev_psi, tors_prev_omg, phi):
    """Next residue's N, CA, C from the previous residue's atoms (all (..., 3))."""
    Nn = _place(N, CA, C, _B["c_n"], _A["ca_c_n"], tors_prev_psi)
    CAn = _place(CA, C, Nn, _B["n_ca"], _A["c_n_ca"], tors_prev_omg)
    Cn = _place(C, Nn, CAn, _B["ca_c"], _A["n_ca_c"], phi)
    return Nn, CAn, Cn


def ss_backbone(L, rng, n_cand=96):
    """Protein-like compact backbone built by NeRF from ideal geometry: helices and strands
    joined by loops; each loop is chosen among n_cand random candidates so that the next
    element packs against what is already built (smallest radius of gyration without CA
    clashes).  Returns (xyz (L,4,3) N,CA,C,CB, torsions (L,3) radians)."""
    states = np.deg2rad(np.array([[-140, 153], [-72, 145], [-122, 117], [-82, -14], [-61, -41], [57, 39]], dtype=float))
    tors = np.zeros((L, 3))
    tors[:, 2] = np.pi
    N = np.zeros((1, 3)); CA = np.array([[_B["n_ca"], 0, 0]])
    C = np.array([[_B["n_ca"] - _B["ca_c"] * np.cos(_A["n_ca_c"]), _B["ca_c"] * np.sin(_A["n_ca_c"]), 0]])
    built = [(N[0], CA[0], C[0])]
    tors[0, :2] = states[4]
    i = 1
    while i < L:
        for attempt in range(8):
            helix = rng.random() < 0.6
            n_ss = int(rng.integers(9, 19) if helix else rng.integers(5, 10))
            n_loop = int(rng.integers(3, 7)) if i > 1 else 0
            n_seg = min(n_loop + n_ss, L - i)
            seg = np.zeros((n_cand, n_seg, 2))
            k = rng.integers(0, 6, size=(n_cand, n_seg))
            seg[:] = states[k] + rng.normal(size=(n_cand, n_seg, 2)) * 0.25
            ss = (states[4] if helix else states[0])
            seg[:, n_loop:] = ss + rng.normal(size=(n_cand, max(n_seg - n_loop, 0), 2)) * 0.08
            pN, pCA, pC = (np.repeat(a[None], n_cand, 0) for a in built[-1])
            prev_psi = np.full(n_cand, tors[i - 1, 1])
            cas, atoms = [], []
            for r in range(n_seg):
                pN, pCA, pC = _extend(pN, pCA, pC, prev_psi, np.pi, seg[:, r, 0])
                prev_psi = seg[:, r, 1]
                cas.append(pCA)
                atoms.append((pN, pCA, pC))
            cas = np.stack(cas, 1)                                   # (cand, n_seg, 3)
            old = np.array([b[1] for b in built])                    # (n_old, 3)
            d = np.linalg.norm(cas[:, :, None, :] - old[None, None, :, :], axis=-1)
            d[:, 0, -1] = 10.0                                       # bonded neighbour
            if n_seg > 1:
                d[:, 1, -1] = np.maximum(d[:, 1, -1], 4.5)
            clash = (d < 4.2).sum((1, 2))
            dn = np.linalg.norm(cas[:, :, None, :] - cas[:, None, :, :], axis=-1)   # within the new segment
            sep = np.abs(np.arange(n_seg)[:, None] - np.arange(n_seg)[None, :])
            clash = clash + ((dn < 4.2) & (sep > 2)[None]).sum((1, 2))
            allca = np.concatenate([np.repeat(old[None], n_cand, 0), cas], 1)
            rg = np.sqrt(((allca - allca.mean(1, keepdims=True)) ** 2).sum(-1).mean(1))
            score = rg + 100.0 * clash
            best = int(np.argmin(score))
            if clash[best] == 0:
                break
        for r in range(n_seg):
            built.append(tuple(a[best] for a in atoms[r]))
            tors[i + r, :2] = seg[best, r]
        i += n_seg
    N = np.array([b[0] for b in built]); CA = np.array([b[1] for b in built]); C = np.array([b[2] for b in built])
    bb, cc = CA - N, C - CA
    CB = -0.58273431 * np.cross(bb, cc) + 0.56802827 * bb - 0.54067466 * cc + CA
    return np.stack([N, CA, C, CB], axis=1), tors


def target(L, seed, dense=False, n_domains=1, two_model=False, protein_like=True):
    """Returns (seq, [npz, ...], native_xyz (L,4,3) N,CA,C,CB)."""
    rng = np.random.default_rng(seed)
    if protein_like:
        xyz, _ = ss_backbone(L, rng)
    else:
        xyz = backbone_from_ca(ca_trace(L, rng, n_domains))
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    seq = "".join(rng.choice(aa, size=L))
    npzs = [distograms(xyz, rng, dense)]
    if two_model:
        rng2 = np.random.default_rng(seed + 1)
        npzs.append(distograms(hinge(xyz, rng2), rng2, dense))
    return seq, npzs, xyz


def random_backbones(N, L, seed):
    """N generic decoy coordinate sets (N, L, 3, 3) [N,CA,CB]: compact random walks with
    atoms hung off the CA trace -- generic geometry for kernel parity tests."""
    rng = np.random.default_rng(seed)
    ca = np.cumsum(rng.normal(size=(N, L, 3)) * 2.2, axis=1)
    n_at = ca + rng.normal(size=(N, L, 3)) * 0.8 + np.array([1.2, 0, 0])
    cb = ca + rng.normal(size=(N, L, 3)) * 0.8 + np.array([0, 1.3, 0])
    return np.stack([n_at, ca, cb], axis=2)
