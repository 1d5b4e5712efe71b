"""Pins oracle/restraints_oracle.c: spline algebra vs scipy, dihedral/angle vs the
reference's numpy geometry (golden), analytic gradient vs central differences."""
import ctypes as C
import numpy as np
import pytest
from scipy.interpolate import CubicSpline

from oracle.tables_oracle import gen_rst_oracle, select_oracle
from oracle import restraints_oracle as ro


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_spline_matches_scipy_clamped(golden_dir):
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    rng = np.random.default_rng(0)
    for name in ro.TYPES:
        rec = rst[name]
        for rule in ("H1", "H2"):
            x, y = ro.apply_end_rule(rec["x"], rec["y"][:40], rec["bin_size"], rule)
            y2 = ro.spline_fit(x, y)
            for r in range(0, 40, 5):
                cs = CubicSpline(x, y[r], bc_type=((1, 0.0), (1, 0.0)))
                for xq in rng.uniform(x[0], x[-1], 25):
                    f, df = ro.splinefunc(x, y[r], y2[r], xq)
                    assert abs(f - cs(xq)) < 1e-11 * max(1, abs(f))
                    assert abs(df - cs(xq, 1)) < 1e-10 * max(1, abs(df))
                # flat outside, knot values hit exactly
                assert ro.splinefunc(x, y[r], y2[r], x[0] - 1.0) == (y[r, 0], 0.0)
                assert ro.splinefunc(x, y[r], y2[r], x[-1] + 3.0) == (y[r, -1], 0.0)
                f, _ = ro.splinefunc(x, y[r], y2[r], x[5])
                assert abs(f - y[r, 5]) < 1e-12


def test_end_rule_shapes(golden_dir):
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    x, y = ro.apply_end_rule(rst["dist"]["x"], rst["dist"]["y"][:3], rst["dist"]["bin_size"], "H1")
    assert len(x) == 37 and x[0] == -0.5 and x[-1] == 20.25 and y.shape == (3, 37)
    assert np.all(y[:, 0] == y[:, 1]) and np.all(y[:, -1] == y[:, -2])
    x, y = ro.apply_end_rule(rst["phi"]["x"], rst["phi"]["y"][:3], rst["phi"]["bin_size"], "H2")
    assert len(x) == 16


def test_geometry_matches_reference_numpy(golden_dir):
    g = np.load(f"{golden_dir}/geometry_random.npz")
    pts = np.ascontiguousarray(g["pts"])
    lib = ro.lib()
    for k in range(pts.shape[1]):
        p = [np.ascontiguousarray(pts[i, k]) for i in range(4)]
        d = lib.trxo_dihedral(_pd(p[0]), _pd(p[1]), _pd(p[2]), _pd(p[3]))
        a = lib.trxo_angle(_pd(p[0]), _pd(p[1]), _pd(p[2]))
        assert abs(d - g["dihedral"][k]) < 1e-13
        assert abs(a - g["angle"][k]) < 1e-13


def _random_backbone(L, rng):
    # a compact random walk of CA with N and CB hung nearby: generic geometry
    ca = np.cumsum(rng.normal(size=(L, 3)) * 2.2, axis=0)
    xyz = np.stack([ca + rng.normal(size=(L, 3)) * 0.8 + [1.2, 0, 0], ca,
                    ca + rng.normal(size=(L, 3)) * 0.8 + [0, 1.3, 0]], axis=1)
    return xyz


@pytest.mark.parametrize("rule", ["H1", "H2"])
def test_gradient_matches_central_differences(golden_dir, rule):
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    sel = select_oracle(rst, 1, 90, 0.05)
    rs = ro.RestraintSetOracle(rst, sel, rule)
    rng = np.random.default_rng(1)
    xyz = _random_backbone(90, rng)
    w = np.array([5.0, 4.0, 4.0])
    E, g = rs.energy_grad(xyz, w)
    assert np.all(np.isfinite(E)) and np.all(np.isfinite(g))
    h = 1e-5
    idx = rng.choice(90 * 9, 60, replace=False)
    for q in idx:
        xp, xm = xyz.copy().reshape(-1), xyz.copy().reshape(-1)
        xp[q] += h; xm[q] -= h
        Ep, _ = rs.energy_grad(xp.reshape(90, 3, 3), w)
        Em, _ = rs.energy_grad(xm.reshape(90, 3, 3), w)
        fd = (Ep @ w - Em @ w) / (2 * h)
        assert abs(fd - g.reshape(-1)[q]) < 2e-5 * max(1.0, abs(fd)), (q, fd, g.reshape(-1)[q])


def test_energy_terms_sum_of_restraints(golden_dir):
    # per-term energies are sums of per-restraint SplineFunc values
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    rs = ro.RestraintSetOracle(rst, select_oracle(rst, 1, 90, 0.05), "H1")
    xyz = _random_backbone(90, np.random.default_rng(2))
    E, _ = rs.energy_grad(xyz)
    s = rs.sets["dist"]
    d = np.linalg.norm(xyz[s["a"], 2] - xyz[s["b"], 2], axis=-1)
    e = sum(ro.splinefunc(s["x"], s["y"][r], s["y2"][r], d[r])[0] for r in range(len(d)))
    assert abs(e - E[0]) < 1e-9 * abs(e)


def test_reference_decoys_sit_in_the_minimum_of_our_restraint_score(golden_dir):
    """An indirect pin against Rosetta's output.  The reference's 8 example decoys were minimised by PyRosetta
    against the restraints of the example npz files; scored with THIS oracle's tables, conventions and splines
    they must sit where a minimiser of the same objective ends up: far below a random start and below the
    natives, at the level this repository's own fold reaches, and each decoy family must prefer the
    distograms it was folded from (wrong dihedral signs or atom orders would destroy all of that)."""
    from oracle import fold_oracle as fo
    from oracle.tables_oracle import gen_rst_oracle, select_oracle
    g = np.load(f"{golden_dir}/example_tmscore.npz")
    names = [str(x) for x in g["names"]]
    xyz = np.stack([g["n"], g["ca"], g["cb"]], axis=2)                       # (10, 90, 3 atoms, 3)
    seq = open(f"{golden_dir}/example_seq.fasta").read().split("\n")[1]
    w = np.array([5.0, 4.0, 4.0])
    tot = {}
    for tag in ("NMR", "Xray"):
        rst = gen_rst_oracle(np.load(f"{golden_dir}/example_{tag}.npz"))
        rs = ro.RestraintSetOracle(rst, select_oracle(rst, 1, 90, 0.05), "H1")
        tot[tag] = np.array([rs.energy_grad(x, w)[0] @ w for x in xyz])
        if tag == "NMR":
            F = fo.FoldOracle(rs, seq)
            start = np.array([rs.energy_grad(F.nerf(t)[:, :3], w)[0] @ w for t in fo.random_torsions(4, 90, 1)])
            ours = F.fold(fo.random_torsions(8, 90, 5), fo.reference_schedule(), m=20, nthreads=8)["terms"][:, :3] @ w
    dec, natives = tot["NMR"][2:], tot["NMR"][:2]
    assert dec.max() < -240000 and natives.max() < -190000 and start.min() > -100000
    assert dec.max() < natives.min()                                          # minimised decoys beat the natives on the predicted restraints
    assert abs(np.median(ours) - np.median(dec)) < 0.03 * abs(np.median(dec))  # our minimiser ends at the same level
    fam_a = [names.index(n) for n in ("conf_1_1", "conf_1_2", "conf_2_3", "conf_2_4")]   # apo-like family
    fam_h = [names.index(n) for n in ("conf_1_3", "conf_1_4", "conf_2_1", "conf_2_2")]   # holo-like family
    assert tot["Xray"][fam_a].max() < tot["Xray"][fam_h].min()               # apo-like decoys fit the X-ray-model distograms better
    assert tot["NMR"][fam_h].max() < tot["NMR"][fam_a].min()                 # holo-like decoys fit the NMR-model distograms better
