"""Outer dynamics loop (N1) == the reference's get_neighbors / pros /
process_distribution_with_pred_distribution, on golden vectors generated from the reference
(tests/golden/make_golden.py: make_dynamics_golden)."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import dynamics


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/dynamics_example48.npz")


def test_six_d_and_binning_match_reference(gold):
    seq = str(gold["seq"])
    d6, o6, t6, p6 = dynamics.six_d(gold["n"], gold["ca"], gold["c"], gold["cb"], seq)
    for got, name in ((d6, "d6"), (o6, "o6"), (t6, "t6"), (p6, "p6")):
        np.testing.assert_allclose(got, gold[name], rtol=0, atol=1e-12)
    jd, jo, jt, jp = dynamics.bin_indices(d6, o6, t6, p6)
    for got, name in ((jd, "jd"), (jo, "jo"), (jt, "jt"), (jp, "jp")):
        np.testing.assert_array_equal(got, gold[name])
    assert jp.max() <= 12 and (jp[jt <= 12] == 0).all()      # the phi-from-theta quirk: negative theta -> bin 0


def test_decay_smooth_matches_reference(gold):
    bins = {"dist": gold["jd"], "omega": gold["jo"], "theta": gold["jt"], "phi": gold["jp"]}
    changed = 0
    for k, j in bins.items():
        got = dynamics.process_distribution(gold[f"in_{k}"], j)
        np.testing.assert_array_equal(got, gold[f"proc_{k}"])          # bit-exact float32
        changed += int((got != gold[f"in_{k}"]).any(-1).sum())
    assert changed > 500
    np.testing.assert_array_equal(dynamics.process_distribution(gold["in_dist"], gold["jd"], norm=False), gold["tmp"])


def test_next_npz_and_reliability(gold):
    npz = {k: gold[f"in_{k}"] for k in ("dist", "omega", "theta", "phi")}
    out = dynamics.next_npz(npz, gold["n"], gold["ca"], gold["c"], gold["cb"], str(gold["seq"]))
    for k in ("dist", "omega", "theta", "phi"):
        np.testing.assert_array_equal(out[k], gold[f"proc_{k}"])
    np.testing.assert_array_equal(out["tmp"], gold["tmp"])
    # second iteration starts from 'tmp' for the convergence signal
    out2 = dynamics.next_npz(out, gold["n"], gold["ca"], gold["c"], gold["cb"], str(gold["seq"]))
    assert np.max(np.abs(out2["tmp"] - out["tmp"])) > 0.0
    tors = np.deg2rad(np.array([[[0, 0, 180], [-60, -40, 180], [60, 40, 180], [-120, 130, 180], [0, 0, 180]]], dtype=float))
    assert dynamics.reliability_score(tors)[0] == pytest.approx(2 / 3)


def test_generate_loop_with_a_stub_folder():
    # loop control: n_init decoys, then one per iteration until 'tmp' stops changing or n_max
    L = 24
    rng = np.random.default_rng(0)
    def rand(nb):
        a = rng.dirichlet(np.full(nb, 0.5), size=(L, L)).astype(np.float32)
        return a
    npz0 = dict(dist=rand(37), omega=rand(25), theta=rand(25), phi=rand(13))
    calls = []
    def fold_fn(npz, n):
        calls.append(n)
        ca = np.cumsum(rng.normal(size=(n, L, 3)) * 2.0, axis=1)
        xyz = np.stack([ca + [1.2, 0.3, 0], ca, ca + [0.2, 1.4, 0.3], ca + [-0.5, 0.2, 1.3], ca + [0, 0, 2.0]], axis=2)
        return dict(xyz=xyz.astype(np.float32), tors=rng.uniform(-np.pi, np.pi, size=(n, L, 3)).astype(np.float32))
    decoys = dynamics.generate(fold_fn, npz0, L, n_init=4, n_max=5)
    assert calls[0] == 4 and all(c == 1 for c in calls[1:])
    assert len(decoys) == 4 + len(calls) - 1 and len(calls) - 1 <= 5


def test_generate_uses_the_decoys_own_cb_like_the_reference(gold):
    # the product loop (pipeline.run_single_from_npz -> dynamics.generate) must take the decoy's real CB for
    # every non-Gly residue, as the reference does when it re-reads the PDB (utils.py:145-150): the
    # distograms handed to the second fold are the reference-run golden ones
    seq = str(gold["seq"])
    L = len(seq)
    npz0 = {k: gold[f"in_{k}"] for k in ("dist", "omega", "theta", "phi")}
    seen = []

    def fold_fn(npz, n):
        seen.append(npz)
        xyz = np.stack([gold["n"], gold["ca"], np.nan_to_num(gold["cb"]), gold["c"], gold["c"] + 1.0], axis=1)
        return dict(xyz=np.repeat(xyz[None], n, 0), tors=np.zeros((n, L, 3)))

    dynamics.generate(fold_fn, npz0, L, n_init=2, n_max=1, seq=seq)
    assert len(seen) == 2
    for k in ("dist", "omega", "theta", "phi"):
        np.testing.assert_array_equal(seen[1][k], gold[f"proc_{k}"])
    np.testing.assert_array_equal(seen[1]["tmp"], gold["tmp"])


@pytest.mark.gpu
def test_device_dynamics_step_is_bit_identical(gold):
    """trx_dyn_step (6D geometry, binning, decay, renormalisation, Gaussian smoothing on the device) against the
    reference-run golden vectors and against the numpy path, two consecutive iterations: realised bins equal,
    all five maps bit-identical (float32), convergence signal equal."""
    from trx2dyn import capi
    seq = str(gold["seq"])
    npz = {k: gold[f"in_{k}"] for k in ("dist", "omega", "theta", "phi")}
    ctx = capi.Context(0)
    st = capi.DynState(ctx, npz)
    chg = st.step(gold["n"], gold["ca"], gold["c"], gold["cb"], seq)
    got = st.get(want_bins=True)
    for b, name in zip(got["bins"], ("jd", "jo", "jt", "jp")):
        np.testing.assert_array_equal(b, gold[name])
    for k in ("dist", "omega", "theta", "phi"):
        np.testing.assert_array_equal(got[k], gold[f"proc_{k}"])
    np.testing.assert_array_equal(got["tmp"], gold["tmp"])
    assert chg == float(np.max(np.abs(gold["in_dist"] - gold["tmp"])))
    host1 = dynamics.next_npz(npz, gold["n"], gold["ca"], gold["c"], gold["cb"], seq)
    # a second iteration with another decoy (the same backbone rotated and jittered)
    rng = np.random.default_rng(5)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    jig = lambda a: np.nan_to_num(a) @ R.T + rng.normal(size=a.shape) * 0.4
    n2, ca2, c2, cb2 = jig(gold["n"]), jig(gold["ca"]), jig(gold["c"]), jig(gold["cb"])
    chg2 = st.step(n2, ca2, c2, cb2, seq)
    got2 = st.get()
    host2 = dynamics.next_npz(host1, n2, ca2, c2, cb2, seq)
    for k in ("dist", "omega", "theta", "phi", "tmp"):
        np.testing.assert_array_equal(got2[k], host2[k], err_msg=k)
    assert chg2 == float(np.max(np.abs(host1["tmp"] - host2["tmp"])))
    # distance-only chain (--no-angle)
    st0 = capi.DynState(ctx, {"dist": gold["in_dist"]}, angle=False)
    st0.step(gold["n"], gold["ca"], gold["c"], gold["cb"], seq)
    np.testing.assert_array_equal(st0.get()["dist"], gold["proc_dist"])
    st0.close(); st.close(); ctx.close()


def test_reliability_score_on_the_reference_example_decoys(golden_dir):
    """calculate_reliability_score (utils_trX2dy/utils.py:337-372) on the reference's own 8 example decoys: the
    reference's ramachandran_score of each (tests/golden/make_golden.py: make_reliability_golden) equals
    reliability_score of the backbone torsions, and the best-decoy pick (first maximum, run_inference.py:61-69)
    is the same."""
    g = np.load(f"{golden_dir}/example_reliability.npz")
    bb = g["bb"]
    tors = dynamics.backbone_torsions(bb[:, :, 0], bb[:, :, 1], bb[:, :, 2])
    got = dynamics.reliability_score(tors)
    np.testing.assert_allclose(got, g["score"], rtol=0, atol=1e-15)
    assert int(np.argmax(got)) == int(np.argmax(g["score"])) and 0.9 < got.min() <= got.max() < 1.0
