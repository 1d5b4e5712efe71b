"""Decoy-set metrics.  The reference compares structures with the prebuilt bin/TMscore
(utils_trX2dy/utils.py:514-540) and its own GloCon arithmetic (:543-567); tests/golden/example_tmscore.npz
holds what those give on the reference's 8 example decoys + 2 natives (made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import metrics


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/example_tmscore.npz")


def test_host_tm_score_and_rmsd_match_the_tmscore_binary(gold):
    ca, tm, rm = gold["ca"], gold["tm"], gold["rmsd"]
    assert abs(tm[2, 0] - 0.6594) < 1e-9 and abs(rm[2, 0] - 3.100) < 1e-9        # BASELINE.md section 2, conf_1_1 vs apo
    for i, j in [(2, 0), (2, 1), (0, 1), (4, 3), (9, 5), (1, 7)]:
        assert abs(metrics.tm_score(ca[i], ca[j]) - tm[i, j]) < 2e-3               # heuristic search: 4th decimal may differ
        assert abs(metrics.rmsd(ca[i], ca[j]) - rm[i, j]) < 6e-4                   # the binary prints 3 decimals


@pytest.mark.gpu
def test_gpu_tmscore_rmsd_and_glocon_matrices(gold):
    from trx2dyn import capi
    ctx = capi.Context(0)
    tm, rm = metrics.tmscore_matrix(ctx, gold["ca"])
    off = ~np.eye(10, dtype=bool)
    assert np.abs(tm - gold["tm"])[off].max() < 2e-3                               # vs bin/TMscore on all 90 ordered pairs
    assert np.abs(rm - gold["rmsd"])[off].max() < 6e-4
    assert np.allclose(np.diag(tm), 1.0) and np.allclose(np.diag(rm), 0.0)
    # the same search as the host metric: equal up to which residues sit exactly on a cut-off
    for i, j in [(2, 0), (0, 1), (9, 5)]:
        assert abs(tm[i, j] - metrics.tm_score(gold["ca"][i], gold["ca"][j])) < 1e-3
        assert abs(rm[i, j] - metrics.rmsd(gold["ca"][i], gold["ca"][j])) < 1e-9
    gl = metrics.glocon_matrix(ctx, gold["cb"])
    np.testing.assert_allclose(gl, gold["glocon"], rtol=1e-12, atol=1e-12)         # fp64 on both sides
    # a larger random set: symmetry, zero diagonal, agreement with numpy
    rng = np.random.default_rng(0)
    M, L = 37, 61
    cb = np.cumsum(rng.normal(size=(M, L, 3)) * 2.2, axis=1)
    gl = metrics.glocon_matrix(ctx, cb)
    d = np.linalg.norm(cb[:, :, None] - cb[:, None], axis=-1)
    d[d > 20.0] = 0.0
    iu = np.triu_indices(L, 1)
    dm = d[:, iu[0], iu[1]]
    diff = np.abs(dm[:, None] - dm[None])
    want = np.where(diff > 3.0, diff, 0.0).sum(-1) / len(iu[0])
    np.testing.assert_allclose(gl, want, rtol=1e-12, atol=1e-12)
    tm, rm = metrics.tmscore_matrix(ctx, cb)
    assert np.allclose(rm, rm.T, atol=1e-9) and np.all(tm <= 1.0 + 1e-12) and np.all(tm >= 0)
    for i, j in [(0, 1), (5, 30), (36, 2)]:
        assert abs(rm[i, j] - metrics.rmsd(cb[i], cb[j])) < 1e-9
        assert abs(tm[i, j] - metrics.tm_score(cb[i], cb[j])) < 2e-3
    ctx.close()


def test_kmeans_clusters_follow_the_reference_settings(gold):
    names = [str(n) for n in gold["names"]][2:]                       # the 8 decoys
    cl = metrics.kmeans_clusters(gold["glocon"][2:, 2:], names, n_clusters=2)
    assert sorted(sum(cl.values(), [])) == sorted(names) and len(cl) == 2
    # the example's decoys split into an apo-like and a holo-like family
    fam = {n: k for k, v in cl.items() for n in v}
    assert fam["conf_1_1"] == fam["conf_1_2"] == fam["conf_2_3"] == fam["conf_2_4"]
    assert fam["conf_1_3"] == fam["conf_1_4"] == fam["conf_2_1"] == fam["conf_2_2"] != fam["conf_1_1"]
