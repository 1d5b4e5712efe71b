"""Size-independent properties of the hot path at BASELINE.json's FULL sizes (configs[2]: L=300,
dist+omega+theta+phi, 4096 decoys), where the oracle is too slow to be the checker:
rigid-motion invariance of the restraint energies, zero net force and torque of their gradient,
linearity of the gradient in the term weights, independence of a decoy's result from the batch it
sits in (bit-exact), and -- for the whole fold -- bit-reproducibility and batch-composition invariance.
Plus the degenerate inputs: no restraints at all, a single decoy, the smallest chain."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import capi, sampler, schedule, synth, tables

pytestmark = pytest.mark.gpu
L, N = 300, 4096


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def full(ctx):
    seq, npzs, nat = synth.target(L, 300, dense=False, two_model=True)
    params = tables.load_params()
    tb = sampler.build_tables(ctx, npzs[0], seq, params)
    rng = np.random.default_rng(0)
    # half random walks (long distances: flat spline tails), half perturbed natives (every spline branch)
    xyz = synth.random_backbones(N, L, seed=1)
    xyz[N // 2:] = nat[None, :, [0, 1, 3]] + rng.normal(size=(N // 2, L, 3, 3)) * 1.5
    yield seq, npzs, nat, tb, xyz
    tb.close()


def test_full_size_rigid_motion_invariance_and_force_balance(full):
    seq, npzs, nat, tb, xyz = full
    assert sum(tb.info()["counts"]) > 60000
    w = np.array([5.0, 4.0, 4.0])
    E, g = tb.energy_grad(xyz, w, capi.F32)
    assert np.all(np.isfinite(E)) and np.all(np.isfinite(g))
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    t = np.array([13.0, -7.0, 21.0])
    moved = (xyz @ Q.T + t).astype(np.float32)
    E2, g2 = tb.energy_grad(moved, w, capi.F32)
    c = np.array(tb.info()["counts"], dtype=float)
    scale = np.maximum(np.abs(E), np.array([c[0], c[1] + c[2], c[3]]))
    # energies depend on internal geometry only (fp32 coordinates after the motion: 1e-5 of the term)
    assert np.all(np.abs(E2 - E) <= 2e-5 * scale)
    gmax = np.abs(g).reshape(N, -1).max(axis=1)
    # the gradient rotates with the frame.  Random walks contain near-degenerate geometry (CA and CB almost on
    # top of each other: dihedral gradients ~ 1/|A|^2 amplify the fp32 rounding of the moved coordinates), so
    # all but a handful of decoys must comply (1.5 A of noise on the natives also produces a few such cases)
    ok = np.abs(g2 - g @ Q.T).reshape(N, -1).max(axis=1) <= 5e-3 * gmax
    assert ok[N // 2:].mean() > 0.995 and ok.mean() > 0.99, (ok[N // 2:].mean(), ok.mean())
    # translation invariance <=> zero net force; rotation invariance <=> zero net torque
    net = g.astype(np.float64).sum(axis=(1, 2))
    tor = np.cross(xyz.astype(np.float64), g.astype(np.float64)).sum(axis=(1, 2))
    assert np.all(np.abs(net).max(axis=1) <= 2e-4 * gmax * np.sqrt(L))
    lever = np.abs(xyz).reshape(N, -1).max(axis=1)
    assert np.all(np.abs(tor).max(axis=1) <= 2e-4 * gmax * lever * np.sqrt(L))


def test_full_size_gradient_is_linear_in_the_weights_and_batch_independent(full):
    seq, npzs, nat, tb, xyz = full
    x = xyz[:1024]
    Ea, ga = tb.energy_grad(x, np.array([1.0, 0.0, 0.0]), capi.F32)
    Eb, gb = tb.energy_grad(x, np.array([0.0, 1.0, 0.0]), capi.F32)
    Ec, gc = tb.energy_grad(x, np.array([0.0, 0.0, 1.0]), capi.F32)
    E, g = tb.energy_grad(x, np.array([5.0, 4.0, 4.0]), capi.F32)
    np.testing.assert_array_equal(Ea, E)                                   # unweighted terms do not depend on w
    np.testing.assert_array_equal(Eb, E)
    gmax = np.abs(g).reshape(len(x), -1).max(axis=1)
    assert np.all(np.abs(5 * ga + 4 * gb + 4 * gc - g).reshape(len(x), -1).max(axis=1) <= 1e-5 * gmax)
    # a decoy's energies and gradient are bit-identical whatever batch it is evaluated in
    pick = np.array([7, 900, 33, 512, 1023, 64])
    Es, gs = tb.energy_grad(x[pick], np.array([5.0, 4.0, 4.0]), capi.F32)
    np.testing.assert_array_equal(Es, E[pick])
    np.testing.assert_array_equal(gs, g[pick])


def test_full_size_fold_is_reproducible_and_batch_independent(ctx, full):
    seq, npzs, nat, tb, xyz = full
    runs = schedule.reference_schedule()
    aa = sampler.aa_index(seq)
    t0 = sampler.random_torsions(256, L, seed=3)
    big = capi.FoldBatch(ctx, [tb], [256], aa, runs)
    a = big.run(t0)
    b = big.run(t0)
    np.testing.assert_array_equal(a["tors"], b["tors"])
    np.testing.assert_array_equal(a["xyz"], b["xyz"])
    big.close()
    small = capi.FoldBatch(ctx, [tb], [64], aa, runs)
    c = small.run(t0[128:192])
    np.testing.assert_array_equal(c["tors"], a["tors"][128:192])
    np.testing.assert_array_equal(c["terms"], a["terms"][128:192])
    small.close()
    w = np.array([5.0, 4.0, 4.0, 1.0, 1.0, 0.5, 0.0, 5.0])
    assert np.median(a["terms"] @ w) < -100000 and np.all(a["iters"] > 50)
    # ... also inside a bench-sized batch (2048 decoys on this table: many more decoy groups per launch, decoys packed
    # and re-packed as they finish); the restraint kernel's summation order must not follow the live-decoy count
    huge = capi.FoldBatch(ctx, [tb], [2048], aa, runs)
    t1 = sampler.random_torsions(2048, L, seed=9)
    t1[1000:1064] = t0[128:192]
    h = huge.run(t1)
    huge.close()
    np.testing.assert_array_equal(h["tors"][1000:1064], c["tors"])
    np.testing.assert_array_equal(h["xyz"][1000:1064], c["xyz"])


def test_degenerate_inputs(ctx):
    # no restraints at all: energies and gradient are exactly zero, and a fold is a pure centroid minimisation
    Ls = 40
    seq, npzs, nat = synth.target(Ls, seed=3)
    empty = {k: (np.zeros(0, np.int32), np.zeros(0, np.int32), np.linspace(0, 1, 5), np.zeros((0, 5))) for k in tables.TYPES}
    tb = capi.Tables(ctx, Ls, empty)
    assert tb.info()["counts"] == [0, 0, 0, 0] and tb.info()["tiles"] == 0
    xyz = synth.random_backbones(5, Ls, seed=2)
    for prec in (capi.F64, capi.F32):
        E, g = tb.energy_grad(xyz, (5.0, 4.0, 4.0), prec)
        assert np.all(E == 0.0) and np.all(g == 0.0)
    batch = capi.FoldBatch(ctx, [tb], [1], sampler.aa_index(seq), schedule.reference_schedule())   # a single decoy
    out = batch.run(sampler.random_torsions(1, Ls, seed=1))
    assert np.all(out["terms"][:, :3] == 0.0) and np.all(np.isfinite(out["xyz"]))
    batch.close(); tb.close()
    # the shortest chain the tables accept, one restraint
    one = {"dist": (np.array([0], np.int32), np.array([1], np.int32), np.array([0.0, 2.0, 4.0, 6.0]), np.array([[3.0, 1.0, 0.0, 0.5]]))}
    tb = capi.Tables(ctx, 2, one)
    x2 = np.zeros((1, 2, 3, 3)); x2[0, 1, :, 0] = 3.0; x2[0, :, 0, 1] = 1.0; x2[0, :, 1, 2] = 1.0
    E, g = tb.energy_grad(x2, (1.0, 1.0, 1.0), capi.F64)
    from oracle import restraints_oracle as ro
    xk, yk = ro.apply_end_rule(one["dist"][2], one["dist"][3], 0.5, "H2")
    f, df = ro.splinefunc(xk, yk[0], ro.spline_fit(xk, yk)[0], 3.0)
    assert abs(E[0, 0] - f) < 1e-12 and abs(g[0, 1, 2, 0] - df) < 1e-12 and abs(g[0, 0, 2, 0] + df) < 1e-12
    tb.close()
    # errors are reported, not swallowed
    with pytest.raises(RuntimeError):
        capi.Tables(ctx, 2, {"dist": (np.array([0], np.int32), np.array([5], np.int32), np.array([0.0, 2.0, 4.0]), np.zeros((1, 3)))})
