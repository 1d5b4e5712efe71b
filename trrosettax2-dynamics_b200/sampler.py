"""Host driver of the fold: initial torsions, batch fold, decoy selection.

Mirrors the reference's per-decoy driver folding/folding.py:main (one process per decoy)
as one batched call: read npz -> gen_rst -> add_rst selection -> tables on device ->
random phi/psi (set_random_dihedral, utils_ros.py:656-696) -> staged minimisation."""
from __future__ import annotations

import numpy as np

from . import capi, schedule, tables

AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
# utils_ros.py:674-696: (phi, psi) states and the cumulative thresholds of random_dihedral
_STATES = np.array([[-140, 153], [-72, 145], [-122, 117], [-82, -14], [-61, -41], [57, 39]], dtype=np.float64)
_EDGES = np.array([0.135, 0.29, 0.363, 0.485, 0.982])


def aa_index(seq, gly_to_ala=True):
    """Residue types for scoring; folding.py:112-115 mutates G->A for the centroid stage."""
    idx = np.array([AA_ORDER.index(c) if c in AA_ORDER else 0 for c in seq], dtype=np.int32)
    if gly_to_ala:
        idx[idx == AA_ORDER.index("G")] = 0
    return idx


def random_torsions(N, L, seed):
    """set_random_dihedral for N decoys: residues 1..L-1 draw from the 6-state table,
    omega = 180; residue L keeps 180/180/180.  (N, L, 3) float32 radians.  The reference
    is unseeded; we take a seed so that runs are reproducible."""
    rng = np.random.default_rng(seed)
    k = np.searchsorted(_EDGES, rng.random((N, L)), side="left")
    t = np.full((N, L, 3), 180.0)
    t[:, :, :2] = _STATES[k]
    t[:, L - 1, :] = 180.0
    return np.deg2rad(t).astype(np.float32)


def build_tables(ctx, npz, seq, params, sep=(1, None), rule="H1", nogly=False, use_orient=None):
    """use_orient: None = what params['USE_ORIENT'] says, else True/False (--orient / --no-orient);
    an npz without orientation maps is folded on distances alone."""
    L = len(seq)
    if use_orient is None:
        use_orient = params.get("USE_ORIENT", True) in (True, "True") and all(k in npz for k in ("omega", "theta", "phi"))
    rst = tables.gen_rst(npz, params, use_orient=use_orient)
    masks = tables.select(rst, sep[0], sep[1] or L, params, seq, nogly)
    return capi.Tables(ctx, L, tables.active_restraints(rst, masks, rule))


def fold(ctx, npzs, seq, n_decoys, seed=0, params=None, runs=None, lbfgs_m=20, rule="H1", max_rounds=20000, reseed=2):
    """Folds n_decoys[t] decoys against npzs[t] (two-model mixing = two entries).
    Returns the dict of capi.FoldBatch.run plus 'model' (index of the npz per decoy)."""
    params = params or tables.load_params()
    runs = runs or schedule.reference_schedule()
    L = len(seq)
    tabs = [build_tables(ctx, npz, seq, params, rule=rule) for npz in npzs]
    batch = capi.FoldBatch(ctx, tabs, n_decoys, aa_index(seq), runs, lbfgs_m)
    out = batch.run(random_torsions(sum(n_decoys), L, seed), max_rounds=max_rounds)
    out["model"] = np.repeat(np.arange(len(npzs)), n_decoys)
    out["status"] = batch.status()
    # failure handling: decoys that report a non-finite energy are folded again from a new random start
    # (the reference has no such step: a failed child process is a missing PDB file)
    for attempt in range(1, reseed + 1):
        bad = np.nonzero(out["status"] & 1)[0]
        if len(bad) == 0:
            break
        fresh = random_torsions(sum(n_decoys), L, seed + 1000003 * attempt)
        start = out["tors"].copy()
        start[bad] = fresh[bad]
        again = batch.run(start, max_rounds=max_rounds)
        st = batch.status()
        for key in ("tors", "xyz", "terms", "evals", "iters"):
            out[key][bad] = again[key][bad]
        out["status"][bad] = st[bad] | 8          # 8: re-seeded
    batch.close()
    for t in tabs:
        t.close()
    return out
