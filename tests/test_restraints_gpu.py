"""GPU parity of K1 (restraint energy + gradient) against the CPU oracle, through the C ABI.

Tolerances (BASELINE.json north_star): fp64 energies 1e-6 relative, gradients 1e-5
relative -- we assert far tighter (1e-10 / 1e-9) because both sides are fp64 and only
summation order and FMA contraction differ.  fp32 (throughput mode) is checked at 2e-6
x max(|E|, #restraints) energy / 2e-3 of the gradient's max norm."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import capi, synth, tables
from oracle import restraints_oracle as ro
from oracle.tables_oracle import gen_rst_oracle, select_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _oracle_and_tables(ctx, npz, L, rule, sep=(1, None), pcut=0.05, use_orient=True):
    sep2 = sep[1] or L
    rst_o = gen_rst_oracle(npz, use_orient=use_orient)
    rs = ro.RestraintSetOracle(rst_o, select_oracle(rst_o, sep[0], sep2, pcut), rule)
    params = tables.load_params()
    params["PCUT"] = pcut
    rst = tables.gen_rst(npz, params, use_orient=use_orient)
    act = tables.active_restraints(rst, tables.select(rst, sep[0], sep2, params), rule)
    return rs, capi.Tables(ctx, L, act)


def _compare(rs, tb, xyz, w, e_tol, g_tol, precision):
    E, g = tb.energy_grad(xyz, w, precision)
    c = tb.info()["counts"]
    for n in range(xyz.shape[0]):
        Eo, go = rs.energy_grad(xyz[n], w)
        # per-term scale: |E| or the number of restraints in the term (terms are sums of
        # O(1) values of both signs, so fp32 error scales with the count, not with |sum|)
        scale = np.maximum(np.abs(Eo), np.array([c[0], c[1] + c[2], c[3]], dtype=float) + 1.0)
        assert np.all(np.abs(E[n] - Eo) <= e_tol * scale), (n, E[n], Eo)
        gmax = np.abs(go).max()
        assert np.abs(g[n] - go).max() <= g_tol * gmax, (n, np.abs(g[n] - go).max(), gmax)


@pytest.mark.parametrize("rule", ["H1", "H2"])
def test_example_fp64_parity(ctx, golden_dir, rule):
    npz = np.load(f"{golden_dir}/example_NMR.npz")
    rs, tb = _oracle_and_tables(ctx, npz, 90, rule)
    assert tb.info()["counts"] == [3226, 2562, 5142, 2541]
    for t, name in enumerate(ro.TYPES):  # spline fit on device == oracle fit
        np.testing.assert_allclose(tb.y2(t), rs.sets[name]["y2"], rtol=1e-11, atol=1e-11)
    xyz = synth.random_backbones(37, 90, seed=5)  # 37: a partial decoy group
    _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 1e-10, 1e-9, capi.F64)
    _compare(rs, tb, xyz[:3], np.array([3.0, 1.0, 1.0]), 1e-10, 1e-9, capi.F64)
    tb.close()


def test_example_fp32_mode(ctx, golden_dir):
    npz = np.load(f"{golden_dir}/example_Xray.npz")
    rs, tb = _oracle_and_tables(ctx, npz, 90, "H1")
    xyz = synth.random_backbones(33, 90, seed=6)
    _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 2e-6, 2e-3, capi.F32)
    tb.close()


@pytest.mark.parametrize("sep", [(1, 12), (12, 24), (24, None), (3, 24)])
def test_separation_windows(ctx, golden_dir, sep):
    # the -m 0 / -m 1 staged schedules of folding.py:125-160 use these windows
    npz = np.load(f"{golden_dir}/example_NMR.npz")
    rs, tb = _oracle_and_tables(ctx, npz, 90, "H1", sep=sep)
    _compare(rs, tb, synth.random_backbones(4, 90, seed=7), np.array([5.0, 4.0, 4.0]), 1e-10, 1e-9, capi.F64)
    tb.close()


def test_no_orient_distance_only(ctx, golden_dir):
    npz = np.load(f"{golden_dir}/example_NMR.npz")
    rs, tb = _oracle_and_tables(ctx, npz, 90, "H1", use_orient=False)
    assert tb.info()["counts"][1:] == [0, 0, 0]
    xyz = synth.random_backbones(5, 90, seed=8)
    _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 1e-10, 1e-9, capi.F64)
    E, _ = tb.energy_grad(xyz)
    assert np.all(E[:, 1:] == 0.0)
    tb.close()


def test_native_like_coordinates_in_range_branch(ctx):
    # decoys near the native exercise the in-range spline branches; random ones the flat tails
    seq, npzs, nat = synth.target(64, seed=11)
    rs, tb = _oracle_and_tables(ctx, npzs[0], 64, "H1")
    rng = np.random.default_rng(0)
    xyz = np.stack([nat[:, [0, 1, 3]] + rng.normal(size=(64, 3, 3)) * s for s in (0.0, 0.05, 0.3, 1.0, 3.0)])
    _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 1e-10, 1e-9, capi.F64)
    # the exact synthetic native holds near-collinear N-CA-CB-CB quadruples whose dihedral
    # gradient (|g| ~ 4e3) is ill-conditioned in fp32: 1e-2 of the max norm there
    _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 2e-6, 1e-2, capi.F32)
    tb.close()


def test_l300_dense_linearity_and_permutation(ctx):
    # full-size properties the oracle is too slow to check decoy by decoy:
    #  (1) E and grad are linear in the weights, (2) a decoy's result does not depend on
    #  which lane / group it sits in, (3) one decoy equals the oracle
    seq, npzs, nat = synth.target(300, seed=300, dense=True)
    params = tables.load_params()
    rst = tables.gen_rst(npzs[0], params)
    act = tables.active_restraints(rst, tables.select(rst, 1, 300, params))
    tb = capi.Tables(ctx, 300, act)
    assert sum(tb.info()["counts"]) > 269000
    xyz = synth.random_backbones(70, 300, seed=9)
    xyz[0] = nat[:, [0, 1, 3]]
    E1, g1 = tb.energy_grad(xyz, (1.0, 0.0, 0.0))
    E2, g2 = tb.energy_grad(xyz, (0.0, 1.0, 0.0))
    E3, g3 = tb.energy_grad(xyz, (0.0, 0.0, 1.0))
    E, g = tb.energy_grad(xyz, (5.0, 4.0, 4.0))
    np.testing.assert_array_equal(E, E1)
    np.testing.assert_allclose(g, 5 * g1 + 4 * g2 + 4 * g3, rtol=1e-10, atol=1e-9)
    perm = np.random.default_rng(1).permutation(70)
    Ep, gp = tb.energy_grad(xyz[perm], (5.0, 4.0, 4.0))
    np.testing.assert_array_equal(Ep, E[perm])          # bit-reproducible regardless of lane
    np.testing.assert_array_equal(gp, g[perm])
    rst_o = gen_rst_oracle(npzs[0])
    rs = ro.RestraintSetOracle(rst_o, select_oracle(rst_o, 1, 300, 0.05), "H1")
    for n in (0, 1):
        Eo, go = rs.energy_grad(xyz[n], (5.0, 4.0, 4.0))
        np.testing.assert_allclose(E[n], Eo, rtol=1e-10)
        assert np.abs(g[n] - go).max() <= 1e-9 * np.abs(go).max()
    tb.close()


def test_bad_arguments_fail_loudly(ctx):
    x = np.array([0.0, 1.0, 2.0, 3.0])
    y = np.zeros((1, 4))
    with pytest.raises(capi.TrxError, match="bad residues"):
        capi.Tables(ctx, 10, {"dist": (np.array([1]), np.array([1]), x, y)})
    with pytest.raises(capi.TrxError, match="duplicate"):
        capi.Tables(ctx, 10, {"dist": (np.array([1, 2]), np.array([2, 1]), x, np.zeros((2, 4)))})
    with pytest.raises(capi.TrxError, match="increasing"):
        capi.Tables(ctx, 10, {"dist": (np.array([1]), np.array([2]), x[::-1].copy(), y)})
    tb = capi.Tables(ctx, 10, {})  # empty restraint set is legal: zero energy, zero gradient
    E, g = tb.energy_grad(synth.random_backbones(2, 10, 0))
    assert np.all(E == 0) and np.all(g == 0)


@pytest.mark.parametrize("rule", ["H1", "H2"])
def test_af2_variant_ca_ca_parity(ctx, golden_dir, rule):
    """-r af2 (utils_ros.py:148-194): 'AtomPair CA a CA b' restraints with 60 (+2) knots on the uneven
    0 / 2.325 / 3.575 / 3.875+0.3125k grid; fp64 and fp32 against the oracle."""
    from oracle.tables_oracle import gen_rst_af2_oracle
    g = np.load(f"{golden_dir}/gen_rst_variants24.npz")
    af2 = dict(dist=g["in_af2_dist"], bins=g["in_af2_bins"])
    L = af2["dist"].shape[0]
    rst_o = gen_rst_af2_oracle(af2)
    rs = ro.RestraintSetOracle(rst_o, select_oracle(rst_o, 1, L, 0.05), rule)
    params = tables.load_params()
    rst = tables.gen_rst(af2, params, False, "af2")
    tb = capi.Tables(ctx, L, tables.active_restraints(rst, tables.select(rst, 1, L, params), rule))
    assert tb.info()["counts"][1:] == [0, 0, 0] and tb.K[0] == (62 if rule == "H1" else 60)
    xyz = synth.random_backbones(40, L, seed=3) * 0.6          # compact: many CA-CA distances inside the knot range
    w = np.array([5.0, 4.0, 4.0])
    _compare(rs, tb, xyz, w, 1e-10, 1e-9, capi.F64)
    _compare(rs, tb, xyz, w, 2e-6, 2e-3, capi.F32)
    E, grad = tb.energy_grad(xyz, w, capi.F64)
    assert np.abs(grad[:, :, 1]).max() > 0 and np.all(grad[:, :, 0] == 0) and np.all(grad[:, :, 2] == 0)   # forces on CA only
    # CA-CA cannot be combined with angular restraints
    full = tables.gen_rst(np.load(f"{golden_dir}/example_NMR.npz"), params)
    act = tables.active_restraints(full, None)
    with pytest.raises(RuntimeError, match="angular"):
        capi.Tables(ctx, 90, act, dist_atom="CA")
    tb.close()


def test_idp_and_gpcr_tables_on_device(ctx, golden_dir):
    """The idp / gpcr variants only change knot values: same kernel, parity against the oracle's tables."""
    from oracle.tables_oracle import gen_idp_rst_oracle, gen_gpcr_rst_oracle
    g = np.load(f"{golden_dir}/gen_rst_variants24.npz")
    r = np.load(f"{golden_dir}/gen_rst_random24.npz")
    inp = {k: r[f"in_{k}"] for k in tables.TYPES}
    inp["idr"] = g["in_idr"]
    known = {k: g[f"in_known_{k}"] for k in ("dist", "omega", "theta_asym", "phi_asym")}
    L = 24
    params = tables.load_params()
    xyz = synth.random_backbones(33, L, seed=8)
    for variant, rst_o in (("idp", gen_idp_rst_oracle(inp, True)), ("gpcr", gen_gpcr_rst_oracle(inp, known, True))):
        rs = ro.RestraintSetOracle(rst_o, select_oracle(rst_o, 1, L, 0.05), "H1")
        rst = tables.gen_rst(inp, params, True, variant, known)
        tb = capi.Tables(ctx, L, tables.active_restraints(rst, tables.select(rst, 1, L, params), "H1"))
        _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 1e-10, 1e-9, capi.F64)
        # fp32: these Dirichlet tables are far rougher than network output (slopes of tens of units per
        # radian), so the 1e-6 rad that fp32 coordinates cost an angle shows up at 1e-5 of the term
        _compare(rs, tb, xyz, np.array([5.0, 4.0, 4.0]), 2e-5, 5e-3, capi.F32)
        tb.close()
