"""The outer 'dynamics' loop of run_inference.py on in-memory arrays.

Reference: after every fold, the decoy is converted back into one-hot 6D geometry, the
distogram bins the decoy realised are decayed where the prediction was not confident, the
result is renormalised and smoothed, and the next fold runs on the new distogram
(run_inference.py:16-143; utils_trX2dy/utils.py:125-249 get_neighbors/pros,
:379-403 process_distribution_with_pred_distribution, :406-476 get_npz_from_pred_pdb).
The reference round-trips through PDB text, Bio.PDB, npz files and 10^4 restraint files per
iteration; here the same arithmetic runs vectorised on the decoy's coordinates.  Quirks kept:
phi is binned from the THETA values (utils.py:226), the decay skips the last bin
(fw = -1, :392), the convergence test uses the un-normalised 'tmp' (run_inference.py:135)."""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter1d

# params("0HD") of utils.py:331-332: backward, forward, P, pcut, decay_rate
P_CONF, P_CUT, DECAY = 0.5, 0.05, 0.50


def _dihedrals(a, b, c, d):
    b0 = a - b
    b1 = c - b
    b2 = d - c
    b1 = b1 / np.linalg.norm(b1, axis=-1, keepdims=True)
    v = b0 - np.sum(b0 * b1, axis=-1, keepdims=True) * b1
    w = b2 - np.sum(b2 * b1, axis=-1, keepdims=True) * b1
    return np.arctan2(np.sum(np.cross(b1, v) * w, axis=-1), np.sum(v * w, axis=-1))


def _angles(a, b, c):
    v = a - b
    v = v / np.linalg.norm(v, axis=-1, keepdims=True)
    w = c - b
    w = w / np.linalg.norm(w, axis=-1, keepdims=True)
    return np.arccos(np.sum(v * w, axis=-1))


def six_d(n, ca, c, cb=None, seq=None, dmax=20.0):
    """get_neighbors (utils.py:125-182): CB-CB distance, omega, theta, phi matrices, zero for
    pairs farther than dmax.  cb: explicit CB where the residue is not Gly (virtual otherwise)."""
    b, cc = ca - n, c - ca
    vcb = -0.58273431 * np.cross(b, cc) + 0.56802827 * b - 0.54067466 * cc + ca
    if cb is not None and seq is not None:
        use = np.array([s != "G" for s in seq]) & ~np.isnan(cb).any(axis=1)
        vcb = np.where(use[:, None], cb, vcb)
    L = len(ca)
    d = np.linalg.norm(vcb[:, None] - vcb[None], axis=-1)
    near = (d <= dmax) & ~np.eye(L, dtype=bool)
    i, j = np.nonzero(near)
    dist6d = np.zeros((L, L)); omega6d = np.zeros((L, L)); theta6d = np.zeros((L, L)); phi6d = np.zeros((L, L))
    dist6d[i, j] = np.linalg.norm(vcb[j] - vcb[i], axis=-1)
    omega6d[i, j] = _dihedrals(ca[i], vcb[i], vcb[j], ca[j])
    theta6d[i, j] = _dihedrals(n[i], ca[i], vcb[i], vcb[j])
    phi6d[i, j] = _angles(ca[i], vcb[i], vcb[j])
    return dist6d, omega6d, theta6d, phi6d


def bin_indices(dist, omega, theta, phi):
    """pros (utils.py:185-249) as bin indices instead of one-hot rows.  Returns jd, jo, jt, jp."""
    jd = (np.arange(2, 20.5, 0.5)[None, None, :] < dist[..., None]).sum(-1)
    jd = np.where(jd >= 37, 0, jd)
    edges = np.arange(-np.pi, np.pi, np.pi / 12)
    jo = (edges[None, None, :] < omega[..., None]).sum(-1)
    jt = (edges[None, None, :] < theta[..., None]).sum(-1)
    jp = (np.arange(0, np.pi, np.pi / 12)[None, None, :] < theta[..., None]).sum(-1)   # sic: theta (utils.py:226)
    gone = jd == 0
    return jd, np.where(gone, 0, jo), np.where(gone, 0, jt), np.where(gone, 0, jp)


def process_distribution(unprocessed, realised_bin, norm=True, smooth=True, sigma=1.0):
    """process_distribution_with_pred_distribution (utils.py:379-403) for all pairs at once.
    realised_bin (L,L): the bin the decoy occupies (argmax of the reference's one-hot)."""
    tmp = np.copy(unprocessed)
    out = np.copy(unprocessed)
    nb = tmp.shape[-1]
    mask = unprocessed.max(axis=-1) < P_CONF
    i, j = np.nonzero(mask)
    k = realised_bin[i, j]
    ok = k <= nb - 2                                   # the slice is empty for the last bin
    ii, jj, kk = i[ok], j[ok], k[ok]
    v = tmp[ii, jj, kk]
    tmp[ii, jj, kk] = np.where(v < P_CUT, v, v * DECAY).astype(tmp.dtype)
    rows = tmp[i, j] / np.sum(tmp[i, j], axis=-1, keepdims=True)
    if smooth:
        rows = gaussian_filter1d(rows, sigma, axis=-1, mode="reflect")
    out[i, j] = rows
    return out if norm else tmp


def next_npz(npz, n, ca, c, cb=None, seq=None, sigma=1.0, angle=True):
    """get_npz_from_pred_pdb twice (processed + tmp), run_inference.py:116-133: the distograms
    for the next iteration from the current ones and the decoy just folded."""
    jd, jo, jt, jp = bin_indices(*six_d(n, ca, c, cb, seq))
    out = {"dist": process_distribution(npz["dist"], jd, sigma=sigma)}
    if angle:
        out["omega"] = process_distribution(npz["omega"], jo, sigma=sigma)
        out["theta"] = process_distribution(npz["theta"], jt, sigma=sigma)
        out["phi"] = process_distribution(npz["phi"], jp, sigma=sigma)
    base_tmp = npz["tmp"] if "tmp" in npz else npz["dist"]
    out["tmp"] = process_distribution(base_tmp, jd, norm=False)
    return out


def reliability_score(tors):
    """calculate_reliability_score (utils.py:337-372): fraction of residues with phi in
    [-180, 0] among those PPBuilder gives both phi and psi (all but the termini).
    tors (..., L, 3) radians -> (...,)."""
    phi = np.rad2deg(np.asarray(tors)[..., 1:-1, 0])
    phi = (phi + 180.0) % 360.0 - 180.0
    return np.mean((phi >= -180.0) & (phi <= 0.0), axis=-1)


def backbone_torsions(n, ca, c):
    """phi, psi, omega (L,3) radians from backbone coordinates (..., L, 3) each; undefined ones (phi of the first,
    psi / omega of the last residue) are pi.  What PPBuilder.get_phi_psi_list gives the reference
    (utils_trX2dy/utils.py:337-349), for decoys that exist as coordinates rather than torsions."""
    n, ca, c = (np.asarray(a, dtype=np.float64) for a in (n, ca, c))
    t = np.full(n.shape[:-1] + (3,), np.pi)
    t[..., 1:, 0] = _dihedrals(c[..., :-1, :], n[..., 1:, :], ca[..., 1:, :], c[..., 1:, :])
    t[..., :-1, 1] = _dihedrals(n[..., :-1, :], ca[..., :-1, :], c[..., :-1, :], n[..., 1:, :])
    t[..., :-1, 2] = _dihedrals(ca[..., :-1, :], c[..., :-1, :], n[..., 1:, :], ca[..., 1:, :])
    return t


def generate(fold_fn, initial_npz, L, n_init=10, n_max=300, sigma=1.0, angle=True, on_decoy=None, seq=None, ctx=None):
    """generate_npz_and_pdb (run_inference.py:16-143) in memory.  fold_fn(npz, n) -> dict with
    'xyz' (n,L,5,3) [N,CA,CB,C,O] and 'tors' (n,L,3): folds n decoys on the given distograms.
    seq: the target's sequence.  The reference re-reads every decoy from its PDB file and takes the
    file's CB for every non-Gly residue (utils.py:145-150), the virtual CB only for Gly; with seq
    given the decoy's own CB is used the same way (without it every CB is the virtual one).
    ctx: a capi.Context -- the distograms then live on the device and the update (6D geometry, binning, decay,
    renormalisation, smoothing) runs there (trx_dyn_step; bit-identical maps); without it the numpy path below.
    Returns the list of decoys [(xyz, tors), ...]: the n_init initial ones, then one per iteration."""
    decoys = []
    out = fold_fn(initial_npz, n_init)
    for k in range(n_init):
        decoys.append((out["xyz"][k], out["tors"][k]))
        if on_decoy:
            on_decoy("initial%d" % k, out["xyz"][k])
    best = int(np.argmax(reliability_score(out["tors"])))     # first maximum, as the reference's loop
    xyz = out["xyz"][best].astype(np.float64)
    state = None
    if ctx is not None:
        from . import capi
        state = capi.DynState(ctx, initial_npz, angle=angle)

    def advance(cur_npz, xyz):
        """-> (next distograms, max change of 'tmp')"""
        if state is not None:
            chg = state.step(xyz[:, 0], xyz[:, 1], xyz[:, 3], cb=xyz[:, 2] if seq else None, seq=seq, sigma=sigma)
            return state.get(), chg
        nxt = next_npz(cur_npz, xyz[:, 0], xyz[:, 1], xyz[:, 3], cb=xyz[:, 2] if seq else None, seq=seq, sigma=sigma, angle=angle)
        base = cur_npz["tmp"] if "tmp" in cur_npz else cur_npz["dist"]
        return nxt, float(np.max(np.abs(np.asarray(base) - nxt["tmp"])))

    cur, _ = advance(initial_npz, xyz)
    it = 0
    while True:
        it += 1
        o = fold_fn(cur, 1)
        decoys.append((o["xyz"][0], o["tors"][0]))
        if on_decoy:
            on_decoy("iter%d" % it, o["xyz"][0])
        if it >= n_max:
            break
        xyz = o["xyz"][0].astype(np.float64)
        cur, chg = advance(cur, xyz)
        if chg < 0.01:
            break
    if state is not None:
        state.close()
    return decoys
