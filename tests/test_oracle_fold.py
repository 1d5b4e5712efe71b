"""Pins oracle/fold_oracle.c on the CPU: NeRF geometry, analytic torsion gradient of the full
score vs central differences, and that the L-BFGS schedule actually folds a small target."""
import ctypes as C
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import metrics, synth
from oracle import fold_oracle as fo, restraints_oracle as ro
from oracle.tables_oracle import gen_rst_oracle, select_oracle


@pytest.fixture(scope="module")
def small():
    L = 40
    seq, npzs, nat = synth.target(L, seed=3)
    rst = gen_rst_oracle(npzs[0])
    rs = ro.RestraintSetOracle(rst, select_oracle(rst, 1, L, 0.05), "H1")
    return seq, nat, fo.FoldOracle(rs, seq)


def test_nerf_ideal_geometry_and_torsions(small):
    seq, nat, F = small
    tors = fo.random_torsions(1, len(seq), 0)[0] + np.random.default_rng(0).normal(size=(len(seq), 3)) * 0.2
    xyz = F.nerf(tors)
    N, CA, CB, Cc, O = (xyz[:, k] for k in range(5))
    assert np.allclose(np.linalg.norm(CA - N, axis=1), 1.458, atol=1e-9)
    assert np.allclose(np.linalg.norm(Cc - CA, axis=1), 1.523, atol=1e-9)
    assert np.allclose(np.linalg.norm(N[1:] - Cc[:-1], axis=1), 1.329, atol=1e-9)
    assert np.allclose(np.linalg.norm(O - Cc, axis=1), 1.231, atol=1e-9)
    lib = ro.lib()
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))
    wrap = lambda a: (a + np.pi) % (2 * np.pi) - np.pi
    for i in (1, 7, 20, 38):
        assert abs(wrap(lib.trxo_dihedral(P(Cc[i - 1]), P(N[i]), P(CA[i]), P(Cc[i])) - tors[i, 0])) < 1e-9      # phi
        assert abs(wrap(lib.trxo_dihedral(P(N[i]), P(CA[i]), P(Cc[i]), P(N[i + 1])) - tors[i, 1])) < 1e-9        # psi
        assert abs(wrap(lib.trxo_dihedral(P(CA[i]), P(Cc[i]), P(N[i + 1]), P(CA[i + 1])) - tors[i, 2])) < 1e-9   # omega
    # CB by the reference's virtual-CB formula (utils_trX2dy/utils.py:132-135)
    b, c = CA - N, Cc - CA
    np.testing.assert_allclose(CB, -0.58273431 * np.cross(b, c) + 0.56802827 * b - 0.54067466 * c + CA, atol=1e-12)


@pytest.mark.parametrize("w", [(5, 4, 4, 1, 1, 0.5), (0, 0, 0, 1, 1, 0), (3, 1, 1, 3, 1, 0.5)])
def test_torsion_gradient_matches_central_differences(small, w):
    seq, nat, F = small
    L = len(seq)
    rng = np.random.default_rng(1)
    t0 = fo.random_torsions(1, L, 2)[0] + rng.normal(size=(L, 3)) * 0.15
    w = np.array(w, dtype=float)
    tot, terms, gt, _ = F.eval(t0, w)
    assert abs(tot - terms @ w) < 1e-9 * max(1.0, abs(tot))
    h = 1e-6
    for (i, k) in [(0, 1), (0, 2), (3, 0), (3, 1), (3, 2), (19, 0), (19, 1), (19, 2), (38, 2), (39, 0), (39, 1)]:
        tp, tm = t0.copy(), t0.copy()
        tp[i, k] += h
        tm[i, k] -= h
        fd = (F.eval(tp, w)[0] - F.eval(tm, w)[0]) / (2 * h)
        assert abs(fd - gt[i, k]) < 1e-5 * max(1.0, abs(fd)), (i, k, fd, gt[i, k])
    assert gt[0, 0] == 0.0 and gt[L - 1, 2] == 0.0


def test_schedule_folds_a_small_target(small):
    seq, nat, F = small
    L = len(seq)
    t0 = fo.random_torsions(6, L, 5)
    out = F.fold(t0, fo.reference_schedule(), m=20, nthreads=6)
    w = np.array([5, 4, 4, 1, 1, 0.5.__float__()])
    e_start = np.array([F.eval(t, w)[0] for t in t0])
    assert np.all(out["terms"] @ w < e_start)
    assert np.all(out["evals"] > out["iters"]) and np.all(out["iters"] > 20)
    tm = np.array([metrics.tm_score(x[:, 1], nat[:, 1]) for x in out["xyz"]])
    assert tm.max() > 0.5
    again = F.fold(t0, fo.reference_schedule(), m=20, nthreads=2)
    np.testing.assert_array_equal(again["tors"], out["tors"])      # deterministic, thread-count independent
