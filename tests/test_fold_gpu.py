"""GPU parity of the fold kernels (NeRF, vdw, rama/omega, torsion gradient, L-BFGS) against
the CPU oracle, through the C ABI.  The fold runs in fp32 on device (fp64 only for energy
accumulation), the oracle in fp64: tolerances are fp32 ones and written beside each check.
The non-restraint terms are approximations of Rosetta's on BOTH sides (same model header);
what is checked here is that the device computes the same function as the oracle and that
decoy quality on the reference's example target matches the oracle's."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import capi, metrics, sampler, schedule, synth, tables
from oracle import fold_oracle as fo, restraints_oracle as ro
from oracle.tables_oracle import gen_rst_oracle, select_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def example(golden_dir):
    seq = open(f"{golden_dir}/example_seq.fasta").read().split("\n")[1]
    npzs = [np.load(f"{golden_dir}/example_NMR.npz"), np.load(f"{golden_dir}/example_Xray.npz")]
    nat = np.load(f"{golden_dir}/example_natives_ca.npz")
    return seq, npzs, nat


def _oracle(npz, seq):
    rst = gen_rst_oracle(npz)
    rs = ro.RestraintSetOracle(rst, select_oracle(rst, 1, len(seq), 0.05), "H1")
    return fo.FoldOracle(rs, seq)


def test_single_evaluation_matches_oracle(ctx, example):
    seq, npzs, _ = example
    L = len(seq)
    params = tables.load_params()
    tb = sampler.build_tables(ctx, npzs[0], seq, params)
    F = _oracle(npzs[0], seq)
    N = 40
    tors = sampler.random_torsions(N, L, seed=3)
    tors[N // 2:] += np.random.default_rng(0).normal(size=(N - N // 2, L, 3)).astype(np.float32) * 0.3
    runs = schedule.reference_schedule()
    batch = capi.FoldBatch(ctx, [tb], [N], sampler.aa_index(seq), runs)
    w = np.array([5.0, 4.0, 4.0, 1.0, 1.0, 0.5, 0.0, 5.0])
    total, terms, gt, xyz = batch.eval(tors, w)
    for n in (0, 7, 20, 33, 39):
        to, termo, gto, xyzo = F.eval(tors[n].astype(np.float64), w)
        # NeRF in fp32 over a 270-atom chain: 2e-3 A
        assert np.abs(xyz[n] - xyzo).max() < 2e-3
        # energies: 1e-5 x max(|E|, #restraints) for sums, looser for the small terms
        assert abs(total[n] - to) < 1e-5 * max(abs(to), 1e4)
        assert np.all(np.abs(terms[n] - termo) < 1e-5 * np.maximum(np.abs(termo), 1e3) + 2e-2)
        # torsion gradient: 2e-3 of its max norm (lever arms amplify fp32 coordinate noise)
        assert np.abs(gt[n] - gto).max() < 2e-3 * np.abs(gto).max()
    batch.close()
    tb.close()


def test_vdw_only_on_clashing_start(ctx):
    # compact random torsions clash heavily: exercises the vdw queue / fixed-point path
    seq, npzs, _ = synth.target(64, seed=5)
    params = tables.load_params()
    tb = sampler.build_tables(ctx, npzs[0], seq, params)
    F = _oracle(npzs[0], seq)
    N = 33
    tors = np.random.default_rng(2).uniform(-np.pi, np.pi, size=(N, 64, 3)).astype(np.float32)
    batch = capi.FoldBatch(ctx, [tb], [N], sampler.aa_index(seq), schedule.reference_schedule())
    w = np.array([0.0, 0.0, 0.0, 1.0, 1.0, 0.5, 0.0, 0.0])
    total, terms, gt, xyz = batch.eval(tors, w)
    assert terms[:, 3].max() > 10.0
    for n in (0, 16, 32):
        to, termo, gto, _ = F.eval(tors[n].astype(np.float64), w)
        assert abs(terms[n, 3] - termo[3]) < 2e-4 * max(termo[3], 1.0)
        assert np.abs(gt[n] - gto).max() < 3e-3 * np.abs(gto).max()
    batch.close()
    tb.close()


def test_fold_example_quality_and_reproducibility(ctx, example):
    seq, npzs, nat = example
    L = len(seq)
    out = sampler.fold(ctx, npzs, seq, [32, 32], seed=11)
    assert np.all(np.isfinite(out["terms"])) and np.all(np.isfinite(out["xyz"]))
    # restraint energy far below the random start, chain connected
    assert np.median(out["terms"][:, 0]) < -15000
    ca = out["xyz"][:, :, 1].astype(np.float64)
    bond = np.linalg.norm(ca[:, 1:] - ca[:, :-1], axis=-1)
    assert abs(bond.mean() - 3.80) < 0.12   # omega (a free DOF under a weak tether) bends CA-CA below 3.80
    tm = np.array([max(metrics.tm_score(c, nat["apo"]), metrics.tm_score(c, nat["holo"])) for c in ca])
    # the reference's own 8 decoys reach TM 0.60-0.67 against the closer native;
    # the CPU oracle of this schedule gives ~0.60 for 7 of 8 starts
    assert np.median(tm) > 0.52, np.sort(tm)
    assert (tm > 0.5).mean() > 0.6
    # oracle, same starts for 6 decoys: same distribution (not same trajectories)
    F = _oracle(npzs[0], seq)
    o = F.fold(sampler.random_torsions(64, L, 11)[:6].astype(np.float64), fo.reference_schedule(), m=20, nthreads=6)
    tmo = np.array([max(metrics.tm_score(c, nat["apo"]), metrics.tm_score(c, nat["holo"])) for c in o["xyz"][:, :, 1]])
    assert abs(np.median(tm[:32]) - np.median(tmo)) < 0.1
    assert abs(np.median(out["terms"][:32, 0]) - np.median(o["terms"][:, 0])) < 0.1 * abs(np.median(o["terms"][:, 0]))
    # bit-reproducible, and independent of batch composition (decoy-sharding invariance)
    again = sampler.fold(ctx, npzs, seq, [32, 32], seed=11)
    np.testing.assert_array_equal(out["tors"], again["tors"])
    tb = sampler.build_tables(ctx, npzs[1], seq, tables.load_params())
    batch = capi.FoldBatch(ctx, [tb], [32], sampler.aa_index(seq), schedule.reference_schedule())
    alone = batch.run(sampler.random_torsions(64, L, 11)[32:])
    np.testing.assert_array_equal(alone["tors"], out["tors"][32:])
    batch.close()
    tb.close()


def test_cartesian_evaluation_matches_oracle(ctx, example):
    """min_mover_cart's objective (folding.py:100-102, scorefxn_cart.wts): restraint + vdw gradients on
    the coordinates, cart_bonded springs, rama / omega from coordinates; fp32 device vs fp64 oracle."""
    seq, npzs, _ = example
    L = len(seq)
    tb = sampler.build_tables(ctx, npzs[0], seq, tables.load_params())
    F = _oracle(npzs[0], seq)
    N = 40
    tors = sampler.random_torsions(N, L, seed=3).astype(np.float64)
    tors += np.random.default_rng(0).normal(size=(N, L, 3)) * 0.3
    rng = np.random.default_rng(1)
    xyz = np.stack([F.nerf(t) for t in tors])
    xyz[N // 2:] += rng.normal(size=(N - N // 2, L, 5, 3)) * 0.04      # half ideal, half strained
    xyz = xyz.astype(np.float32)
    batch = capi.FoldBatch(ctx, [tb], [N], sampler.aa_index(seq), schedule.reference_schedule())
    for w in (np.array([5.0, 4.0, 4.0, 0.5, 1.0, 0.5, 0.1, 3.0]), np.array([0.0, 0.0, 0.0, 0.0, 1.0, 0.5, 1.0, 1.0])):
        total, terms, grad, back = batch.eval_cart(xyz, w)
        for n in (0, 7, 19, 20, 33, 39):
            to, termo, go = F.eval_cart(xyz[n].astype(np.float64), w)
            assert abs(total[n] - to) < 1e-5 * max(abs(to), 1e4)
            # springs: (d - d0)^2 with d ~ 1.5 A in fp32 -> 1e-4 relative + 1e-3 absolute
            assert np.all(np.abs(terms[n] - termo) < 1e-4 * np.maximum(np.abs(termo), 1e2) + 2e-2), (terms[n], termo)
            assert np.abs(grad[n] - go).max() < 1e-3 * np.abs(go).max()
            d = (back[n] - F.torsions(xyz[n].astype(np.float64)) + np.pi) % (2 * np.pi) - np.pi
            assert np.abs(d).max() < 2e-4
    # the schedule without a Cartesian run has no Cartesian buffers: the entry refuses loudly
    plain = capi.FoldBatch(ctx, [tb], [N], sampler.aa_index(seq), schedule.reference_schedule(cartesian=False))
    with pytest.raises(RuntimeError, match="Cartesian"):
        plain.eval_cart(xyz, np.ones(7))
    plain.close(); batch.close(); tb.close()


def test_cartesian_stage_in_the_schedule(ctx):
    """Segmented schedule: torsion runs -> Cartesian run -> remove_clash.  (a) reference thresholds: the
    final min_mover1 rebuilds ideal bonds; (b) remove_clash skipped: the decoy keeps (holds) the Cartesian
    coordinates, and what is reported is exactly their score."""
    seq, npzs, nat = synth.target(64, seed=7)
    L = len(seq)
    tb = sampler.build_tables(ctx, npzs[0], seq, tables.load_params())
    rst = gen_rst_oracle(npzs[0])
    F = fo.FoldOracle(ro.RestraintSetOracle(rst, select_oracle(rst, 1, L, 0.05), "H1"), seq)
    aa = sampler.aa_index(seq)
    t0 = sampler.random_torsions(64, L, seed=2)
    runs = schedule.reference_schedule()
    batch = capi.FoldBatch(ctx, [tb], [64], aa, runs)
    out = batch.run(t0)
    bond = np.linalg.norm(out["xyz"][:, :, 1] - out["xyz"][:, :, 0], axis=-1)
    assert np.abs(bond - 1.458).max() < 1e-4 and np.all(out["terms"][:, 6] == 0.0)
    np.testing.assert_array_equal(batch.run(t0)["tors"], out["tors"])             # bit-reproducible
    batch.close()
    torsion_only = capi.FoldBatch(ctx, [tb], [64], aa, schedule.reference_schedule(cartesian=False)[:8])
    base = torsion_only.run(t0)
    torsion_only.close()
    for r in runs[9:]:
        r.clash_thr = 1e9
    batch = capi.FoldBatch(ctx, [tb], [64], aa, runs)
    held = batch.run(t0)
    bond = np.linalg.norm(held["xyz"][:, :, 1] - held["xyz"][:, :, 0], axis=-1)
    # cart_bonded carries weight 0.1 against restraint weights 5/4/4 (scorefxn_cart.wts) and this synthetic
    # target's distograms are sharp: bonds give visibly.  Same distribution as the oracle's Cartesian stage.
    assert np.abs(bond - 1.458).max() > 1e-3 and np.all(held["terms"][:, 6] > 0.0)
    oruns = fo.reference_schedule(cartesian=True)
    for r in oruns[9:]:
        r.clash_thr = 1e9
    o = F.fold(t0[:8].astype(np.float64), oruns, m=20, nthreads=8)
    obond = np.linalg.norm(o["xyz"][:, :, 1] - o["xyz"][:, :, 0], axis=-1)
    assert abs(np.abs(bond - 1.458).mean() - np.abs(obond - 1.458).mean()) < 0.05
    assert abs(np.median(held["terms"][:, 6]) - np.median(o["terms"][:, 6])) < 0.3 * np.median(o["terms"][:, 6])
    wc = np.array(list(runs[8].w))
    assert abs(np.median(held["terms"] @ wc) - np.median(o["terms"] @ wc)) < 0.03 * abs(np.median(o["terms"] @ wc))
    assert np.all(held["evals"] > base["evals"]) and np.all(held["iters"] > base["iters"])
    w = np.array(list(runs[8].w))
    e_start = np.array([F.eval_cart(base["xyz"][n].astype(np.float64), w)[0] for n in range(0, 64, 7)])
    for k, n in enumerate(range(0, 64, 7)):
        to, termo, _ = F.eval_cart(held["xyz"][n].astype(np.float64), w)
        assert np.all(np.abs(held["terms"][n] - termo) < 1e-4 * np.maximum(np.abs(termo), 1e2) + 2e-2)
        d = (held["tors"][n] - F.torsions(held["xyz"][n].astype(np.float64)) + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(d).max() < 2e-4
        assert to < e_start[k]                       # the Cartesian run lowered its own objective
    tm = np.array([metrics.tm_score(c, nat[:, 1]) for c in held["xyz"][:, :, 1]])
    assert np.median(tm) > 0.5
    # independent of batch composition
    half = capi.FoldBatch(ctx, [tb], [32], aa, runs)
    sub = half.run(t0[32:])
    np.testing.assert_array_equal(sub["tors"], held["tors"][32:])
    np.testing.assert_array_equal(sub["xyz"], held["xyz"][32:])
    half.close(); batch.close()
    # a Cartesian segment starts from what the torsion-space segment before it parked: a schedule may not open with one
    with pytest.raises(capi.TrxError, match="may not open with a Cartesian run"):
        capi.FoldBatch(ctx, [tb], [32], aa, [runs[8]])
    tb.close()


def test_decoy_distributions_match_the_oracle(ctx, example):
    """north_star: 'the RMSD/TM-score distribution of final decoys statistically matched'.  The device folds in fp32
    with its own summation orders, so trajectories differ from the fp64 oracle's; the DISTRIBUTIONS over random
    starts must not.  Two-sample Kolmogorov-Smirnov tests, 128 device decoys against 32 oracle decoys of the
    reference's example target: TM-score and RMSD to the closer native, total score."""
    from scipy.stats import ks_2samp
    seq, npzs, nat = example
    L = len(seq)
    out = sampler.fold(ctx, [npzs[0]], seq, [128], seed=21)
    F = _oracle(npzs[0], seq)
    o = F.fold(sampler.random_torsions(32, L, 77).astype(np.float64), fo.reference_schedule(), m=20, nthreads=16)

    def quality(ca):
        tm = np.array([max(metrics.tm_score(c, nat["apo"]), metrics.tm_score(c, nat["holo"])) for c in ca])
        rm = np.array([min(metrics.rmsd(c, nat["apo"]), metrics.rmsd(c, nat["holo"])) for c in ca])
        return tm, rm
    tm_d, rm_d = quality(out["xyz"][:, :, 1].astype(np.float64))
    tm_o, rm_o = quality(o["xyz"][:, :, 1])
    w = np.array([5.0, 4.0, 4.0, 1.0, 1.0, 0.5, 0.0, 5.0])
    for name, a, b in (("TM", tm_d, tm_o), ("RMSD", rm_d, rm_o), ("score", out["terms"] @ w, o["terms"] @ w)):
        p = ks_2samp(a, b).pvalue
        assert p > 0.01, (name, p, np.median(a), np.median(b))
    assert abs(np.median(tm_d) - np.median(tm_o)) < 0.03 and abs(np.median(rm_d) - np.median(rm_o)) < 0.5
    # the work spent is of the same size (fp32 line searches stop a few percent earlier than the fp64 oracle's)
    assert abs(np.median(out["evals"]) - np.median(o["evals"])) < 0.2 * np.median(o["evals"])


def test_packing_of_unfinished_decoys_changes_nothing(ctx, monkeypatch):
    """The device packs the unfinished decoys of a block to the front when they fill less than half of it
    (migration: positions are swapped, results go back to the caller's order).  Bit-identical to a run with the
    packing disabled -- torsions, coordinates, terms, counters, MC acceptance -- for two table blocks, a block
    size that is not a multiple of 32, the Cartesian segment and Monte-Carlo cycles."""
    seq, npzs, nat = synth.target(48, seed=12, two_model=True)
    params = tables.load_params()
    tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
    aa = sampler.aa_index(seq)
    nd = [160, 75]
    t0 = sampler.random_torsions(sum(nd), 48, seed=4)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRX_NO_MIGRATE", flag)
        batch = capi.FoldBatch(ctx, tabs, nd, aa, schedule.reference_schedule())
        res["fold" + flag] = batch.run(t0)
        batch.close()
        batch = capi.FoldBatch(ctx, tabs, nd, aa, schedule.mc_schedule(mc_max_iter=60))
        res["mc" + flag] = batch.run_mc(t0, cycles=3, kT=2.0, sigma_deg=25.0, seed=5)
        batch.close()
    for kind in ("fold", "mc"):
        a, b = res[kind + "1"], res[kind + "0"]
        for key in a:
            if key != "rounds":
                np.testing.assert_array_equal(a[key], b[key], err_msg="%s %s" % (kind, key))
    assert res["mc0"]["accepted"].sum() > 0 and len(set(res["fold0"]["evals"].tolist())) > 20   # decoys do finish at different times
    for t in tabs:
        t.close()


def test_folding_cli_drop_in(tmp_path, golden_dir, example):
    """The exact command utils_trX2dy/utils.py:491-498 builds (+ seed), and the batched form."""
    import os, subprocess, sys
    from trx2dyn import pdbio
    seq, _, nat = example
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "initial0.pdb"
    cmd = [sys.executable, "./folding/folding.py", "-NPZ", f"{golden_dir}/example_NMR.npz", "-FASTA",
           f"{golden_dir}/example_seq.fasta", "-OUT", str(out), "-m", "2", "--orient", "-r", "no-idp", "--seed", "4"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "*** time:" in r.stdout
    s2, at = pdbio.read_backbone(str(out))
    assert s2 == seq and len(s2) == 90
    gly = np.array([c == "G" for c in seq])
    assert np.isnan(at["CB"][gly]).all() and not np.isnan(at["CB"][~gly]).any()
    pep = np.linalg.norm(at["N"][1:] - at["C"][:-1], axis=1)
    assert np.allclose(pep, 1.329, atol=0.01)                     # bonded peptide C-N for PPBuilder
    tm = max(metrics.tm_score(at["CA"], nat["apo"]), metrics.tm_score(at["CA"], nat["holo"]))
    assert tm > 0.3
    # batched: 8 decoys in one launch, staged mode 0, distance-only
    pat = tmp_path / "b" / "initial{i}.pdb"
    cmd = [sys.executable, "./folding/folding.py", "-NPZ", f"{golden_dir}/example_Xray.npz", "-FASTA",
           f"{golden_dir}/example_seq.fasta", "-OUT", str(pat), "-m", "0", "--no-orient", "-r", "no-idp",
           "--ndecoy", "8", "--start-id", "3", "--seed", "1", "--no-fastrelax"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted(os.listdir(tmp_path / "b"))
    assert files == ["initial%d.pdb" % i for i in (10, 3, 4, 5, 6, 7, 8, 9)]


def test_folding_cli_variants(tmp_path, golden_dir):
    """-r idp -m 3 (order / disorder stages, folding.py:173-186), -r gpcr with -KNOWN, -r af2 --no-orient."""
    import os, subprocess, sys
    from trx2dyn import pdbio
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(f"{golden_dir}/gen_rst_variants24.npz")
    r = np.load(f"{golden_dir}/gen_rst_random24.npz")
    L = 24
    seq = "MKTAYIAKQRQISFVKSHFSRQLE"
    (tmp_path / "s.fasta").write_text(">s\n%s\n" % seq)
    np.savez(tmp_path / "idp.npz", idr=g["in_idr"], **{k: r[f"in_{k}"] for k in ("dist", "omega", "theta", "phi")})
    np.savez(tmp_path / "known.npz", **{k: g[f"in_known_{k}"] for k in ("dist", "omega", "theta_asym", "phi_asym")})
    np.savez(tmp_path / "af2.npz", dist=g["in_af2_dist"], bins=g["in_af2_bins"])
    base = [sys.executable, "./folding/folding.py", "-FASTA", str(tmp_path / "s.fasta"), "--seed", "3", "--no-fastrelax"]
    runs = {"idp3": ["-NPZ", str(tmp_path / "idp.npz"), "-r", "idp", "-m", "3"],
            "gpcr": ["-NPZ", str(tmp_path / "idp.npz"), "-r", "gpcr", "-KNOWN", str(tmp_path / "known.npz"), "-m", "1"],
            "af2": ["-NPZ", str(tmp_path / "af2.npz"), "-r", "af2", "--no-orient"]}
    for tag, extra in runs.items():
        out = tmp_path / (tag + ".pdb")
        res = subprocess.run(base + extra + ["-OUT", str(out)], cwd=root, capture_output=True, text=True)
        assert res.returncode == 0, (tag, res.stderr[-2000:])
        s2, at = pdbio.read_backbone(str(out))
        assert s2 == seq and np.isfinite(at["CA"]).all()
        pep = np.linalg.norm(at["N"][1:] - at["C"][:-1], axis=1)
        assert np.all(pep < 1.8)                                      # bonded peptide C-N for PPBuilder
    res = subprocess.run(base + ["-NPZ", str(tmp_path / "af2.npz"), "-r", "af2", "--orient", "-OUT", str(tmp_path / "x.pdb")],
                         cwd=root, capture_output=True, text=True)
    assert res.returncode != 0 and "AF2 Not support" in res.stderr   # the reference's own refusal (utils_ros.py:150)
    res = subprocess.run(base + ["-NPZ", str(tmp_path / "idp.npz"), "-r", "gpcr", "-OUT", str(tmp_path / "x.pdb")],
                         cwd=root, capture_output=True, text=True)
    assert res.returncode != 0 and "-KNOWN" in res.stderr


def test_continuous_batching_is_bit_identical(ctx):
    """trx_fold_run_queue: 235 decoys of two table blocks folded through 96 positions (a position is refilled as
    soon as its decoy leaves the schedule segment in progress) give, decoy for decoy and bit for bit, what one
    resident batch of 235 positions gives -- torsions, coordinates, terms, counters, MC acceptance -- through the
    torsion | Cartesian | torsion segments and through Monte-Carlo cycles.  Also a queue shorter than the batch."""
    seq, npzs, nat = synth.target(48, seed=12, two_model=True)
    params = tables.load_params()
    tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
    aa = sampler.aa_index(seq)
    nd = [160, 75]
    t0 = sampler.random_torsions(sum(nd), 48, seed=4)
    big = capi.FoldBatch(ctx, tabs, nd, aa, schedule.reference_schedule())
    ref = big.run(t0)
    k1_ref = big.k1_evals()
    big.close()
    small = capi.FoldBatch(ctx, tabs, [64, 32], aa, schedule.reference_schedule())
    got = small.run_queue(t0, nd)
    for key in ref:
        if key != "rounds":
            np.testing.assert_array_equal(got[key], ref[key], err_msg=key)
    assert got["rounds"] > ref["rounds"]                      # fewer positions, more rounds
    assert small.k1_evals() == k1_ref                         # the same restraint-kernel work, decoy for decoy
    assert 0 < k1_ref[0] < ref["evals"][:160].sum() + 3 * 160  # vdw-only evaluations skip the kernel; 3 closing evaluations per decoy
    few = small.run_queue(np.concatenate([t0[:20], t0[160:165]]), [20, 5])
    np.testing.assert_array_equal(few["tors"], np.concatenate([ref["tors"][:20], ref["tors"][160:165]]))
    np.testing.assert_array_equal(few["xyz"], np.concatenate([ref["xyz"][:20], ref["xyz"][160:165]]))
    small.close()
    runs = schedule.mc_schedule(mc_max_iter=60)
    big = capi.FoldBatch(ctx, tabs, nd, aa, runs)
    ref = big.run_mc(t0, cycles=3, kT=2.0, sigma_deg=25.0, seed=5)
    big.close()
    small = capi.FoldBatch(ctx, tabs, [64, 32], aa, runs)
    got = small.run_mc(t0, cycles=3, kT=2.0, sigma_deg=25.0, seed=5, nq=nd)
    for key in ref:
        if key != "rounds":
            np.testing.assert_array_equal(got[key], ref[key], err_msg="mc " + key)
    assert ref["accepted"].sum() > 0
    small.close()
    for t in tabs:
        t.close()


def test_round_budget_closes_every_decoy(ctx):
    """max_rounds exhausted mid-schedule: every decoy (running, or still waiting in the queue) is closed at its
    accepted point with one consistent evaluation -- finite coordinates whose NeRF rebuild is the torsions
    returned, terms of those coordinates."""
    seq, npzs, nat = synth.target(40, seed=3)
    params = tables.load_params()
    tb = sampler.build_tables(ctx, npzs[0], seq, params)
    runs = schedule.reference_schedule()
    batch = capi.FoldBatch(ctx, [tb], [32], sampler.aa_index(seq), runs)
    t0 = sampler.random_torsions(80, 40, seed=1)
    out = batch.run_queue(t0, [80], max_rounds=48)
    assert np.all(np.isfinite(out["xyz"])) and np.all(np.isfinite(out["terms"]))
    assert np.all(batch.status() & 4) and not np.any(batch.status() & 1)    # TRX_DECOY_UNFINISHED for every decoy, none non-finite
    assert out["evals"][:32].min() > 0 and out["evals"][64:].max() == 0      # the tail of the queue never started
    np.testing.assert_array_equal(out["tors"][64:], t0[64:])
    chk = capi.FoldBatch(ctx, [tb], [80], sampler.aa_index(seq), runs)
    w = np.array(list(runs[5].w))
    _, terms, _, xyz = chk.eval(out["tors"], w)
    assert np.abs(xyz - out["xyz"]).max() < 1e-4
    assert np.allclose(terms, out["terms"], rtol=1e-6, atol=1e-6)
    full = batch.run_queue(t0, [80])
    assert not np.any(batch.status() & (1 | 4))                             # finished, finite
    assert np.all((full["terms"] @ w) < (out["terms"] @ w)[:80] + 1e-6)
    chk.close(); batch.close(); tb.close()


def test_pair_list_changes_nothing(ctx, monkeypatch):
    """The vdw / hydrogen-bond pair search keeps a Verlet list per position (partners within reach + a 2 A skin,
    rebuilt when a residue has used up half the skin).  Energies are summed as 64-bit and gradients as 32-bit
    fixed point, so an evaluation through the list is the same bits as one through the full scan: the whole fold
    (two table blocks, Cartesian segment, Monte-Carlo cycles, migration) is bit-identical with TRX_NO_NBL=1."""
    seq, npzs, nat = synth.target(56, seed=21, two_model=True)
    params = tables.load_params()
    tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
    aa = sampler.aa_index(seq)
    nd = [96, 45]
    t0 = sampler.random_torsions(sum(nd), 56, seed=8)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRX_NO_NBL", flag)
        batch = capi.FoldBatch(ctx, tabs, nd, aa, schedule.reference_schedule())
        res["fold" + flag] = batch.run(t0)
        batch.close()
        batch = capi.FoldBatch(ctx, tabs, nd, aa, schedule.mc_schedule(mc_max_iter=60))
        res["mc" + flag] = batch.run_mc(t0, cycles=2, kT=2.0, sigma_deg=25.0, seed=5)
        batch.close()
    for kind in ("fold", "mc"):
        a, b = res[kind + "1"], res[kind + "0"]
        for key in a:
            np.testing.assert_array_equal(a[key], b[key], err_msg="%s %s" % (kind, key))
    for t in tabs:
        t.close()


def test_monte_carlo_extension(ctx):
    """Extension with no reference behaviour: checks the invariants it can have --
    Metropolis never loses the best state at kT -> 0, counters are sane, trajectories are
    reproducible and independent of how decoys are batched (id_offset)."""
    seq, npzs, nat = synth.target(64, seed=7)
    params = tables.load_params()
    tb = sampler.build_tables(ctx, npzs[0], seq, params)
    runs = schedule.mc_schedule(mc_max_iter=100)
    w = np.array(list(runs[-1].w))
    t0 = sampler.random_torsions(64, 64, seed=2)
    batch = capi.FoldBatch(ctx, [tb], [64], sampler.aa_index(seq), runs)
    base = batch.run_mc(t0, cycles=0)
    mc = batch.run_mc(t0, cycles=6, kT=1e-6, sigma_deg=25.0, seed=9)
    e0, e1 = base["terms"] @ w, mc["terms"] @ w
    assert np.all(mc["accepted"] >= 0) and np.all(mc["accepted"] <= 6) and mc["accepted"].sum() > 0
    assert np.all(e1 <= e0 + 1e-3 * np.abs(e0))          # greedy MC (kT ~ 0) cannot end above its start
    assert e1.mean() < e0.mean()
    assert np.all(mc["evals"] > base["evals"])
    hot = batch.run_mc(t0, cycles=6, kT=50.0, sigma_deg=25.0, seed=9)
    assert hot["accepted"].sum() >= mc["accepted"].sum()   # a hotter chain accepts at least as often
    again = batch.run_mc(t0, cycles=6, kT=1e-6, sigma_deg=25.0, seed=9)
    np.testing.assert_array_equal(again["tors"], mc["tors"])
    half = capi.FoldBatch(ctx, [tb], [32], sampler.aa_index(seq), runs)
    sub = half.run_mc(t0[32:], cycles=6, kT=1e-6, sigma_deg=25.0, seed=9, id_offset=32)
    np.testing.assert_array_equal(sub["tors"], mc["tors"][32:])
    np.testing.assert_array_equal(sub["accepted"], mc["accepted"][32:])
    batch.close(); half.close(); tb.close()


def test_dynamics_pipeline_on_example(tmp_path, golden_dir, example, ctx):
    """run_inference's fold -> decay -> fold loop in process (N1/N2): two models, 4 initial decoys
    each, up to 3 iterations; both natives' basins should be visited by some decoy."""
    from trx2dyn import pipeline, pdbio
    seq, _, nat = example
    files = pipeline.run_single_from_npz("seq", f"{golden_dir}/example_seq.fasta",
                                         [f"{golden_dir}/example_NMR.npz", f"{golden_dir}/example_Xray.npz"],
                                         str(tmp_path), init_num=4, n_max=3, seed=5, ctx=ctx)
    names = sorted(os.path.basename(f) for f in files) if False else [f.split("/")[-1] for f in files]
    assert names[:4] == ["conf_1_%d.pdb" % k for k in range(1, 5)]
    n1 = sum(n.startswith("conf_1_") for n in names)
    n2 = sum(n.startswith("conf_2_") for n in names)
    assert 5 <= n1 <= 7 and 5 <= n2 <= 7
    tms = []
    for f in files:
        s2, at = pdbio.read_backbone(f)
        assert s2 == seq
        tms.append((metrics.tm_score(at["CA"], nat["apo"]), metrics.tm_score(at["CA"], nat["holo"])))
    tms = np.array(tms)
    assert tms.max(axis=0).min() > 0.5, tms


def test_recycled_device_blocks_change_nothing(example):
    """Tables, fold batches and their scratch come from the context's block pool (csrc/context.cu): a dynamics loop
    rebuilds them every iteration.  A fold on recycled blocks -- of a DIFFERENT, larger earlier owner -- returns the
    same bits as the same fold on a fresh context."""
    seq, npzs, _ = example
    L = len(seq)
    params = tables.load_params()
    runs = schedule.reference_schedule()
    tors = sampler.random_torsions(24, L, seed=11)

    def fold(c, npz, n):
        tb = sampler.build_tables(c, npz, seq, params)
        batch = capi.FoldBatch(c, [tb], [n], sampler.aa_index(seq), runs)
        out = batch.run(tors[:n])
        batch.close()
        tb.close()
        return out

    fresh = capi.Context(0)
    ref = fold(fresh, npzs[0], 24)
    fresh.close()
    c = capi.Context(0)
    fold(c, npzs[1], 24)          # other restraints, same sizes: its blocks are what the next fold gets
    again = fold(c, npzs[0], 24)
    small = fold(c, npzs[0], 7)   # smaller request served from larger released blocks
    c.close()
    for k in ("tors", "xyz", "terms", "evals"):
        assert np.array_equal(ref[k], again[k]), k
        assert np.array_equal(ref[k][:7], small[k]), k


def test_context_may_be_destroyed_before_its_children(golden_dir):
    """A garbage collector picks its own order (Python at interpreter exit destroys the context first): tables, fold
    batches and dynamics states hold a reference on their context (csrc/context.cu: ctx_release), so either order works.
    Run in a child process: the failure mode is a crash."""
    import subprocess
    import sys
    code = f"""
import numpy as np, ctypes as C
import trx2dyn
from trx2dyn import capi, sampler, schedule, tables
seq = open(r"{golden_dir}/example_seq.fasta").read().split("\\n")[1]
npz = dict(np.load(r"{golden_dir}/example_NMR.npz"))
ctx = capi.Context(0)
tb = sampler.build_tables(ctx, npz, seq, tables.load_params())
batch = capi.FoldBatch(ctx, [tb], [4], sampler.aa_index(seq), schedule.reference_schedule())
state = capi.DynState(ctx, npz)
out = batch.run(sampler.random_torsions(4, len(seq), 0))
capi.lib().trx_ctx_destroy(ctx._h); ctx._h = C.c_void_p()      # the context handle goes first
batch.close(); state.close(); tb.close()                        # ... and its children after it
ctx2 = capi.Context(0)                                           # and: nothing closed at all, left to interpreter exit
tb2 = sampler.build_tables(ctx2, npz, seq, tables.load_params())
state2 = capi.DynState(ctx2, npz)
print("ok", float(out["terms"].sum()))
"""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(golden_dir)))   # the repository root, whatever the cwd
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root,
                       env=dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", "")))
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert r.stdout.strip().startswith("ok")
