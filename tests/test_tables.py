"""Product table builder (trx2dyn.tables) == oracle restatement, value for value."""
import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import tables
from oracle.tables_oracle import gen_rst_oracle, select_oracle, PARAMS
from oracle.restraints_oracle import apply_end_rule


def _inputs(golden_dir):
    yield "example_NMR", np.load(f"{golden_dir}/example_NMR.npz")
    yield "example_Xray", np.load(f"{golden_dir}/example_Xray.npz")
    g = np.load(f"{golden_dir}/gen_rst_random24.npz")
    yield "random24", {k: g[f"in_{k}"] for k in tables.TYPES}


def test_params_file_matches_reference_constants():
    assert tables.load_params() == {**PARAMS, "WDIR": "/dev/shm"}


def test_gen_rst_equals_oracle(golden_dir):
    params = tables.load_params()
    for tag, npz in _inputs(golden_dir):
        for orient in (True, False):
            got = tables.gen_rst(npz, params, use_orient=orient)
            want = gen_rst_oracle(npz, use_orient=orient)
            assert list(got) == list(want)
            for name in got:
                for key in ("a", "b", "p", "x", "y"):
                    np.testing.assert_array_equal(got[name][key], want[name][key], err_msg=f"{tag} {name} {key}")
                assert got[name]["bin_size"] == want[name]["bin_size"]


def test_round_decimals_ties_and_signs():
    v = np.array([0.0005, 0.0015, 2.5e-4, -0.0004, -0.0005, 1.0005, 6.7025, -6.7035, 123.4565, 1e-9])
    for nd in (3, 5):
        want = np.array([float(("%%.%df" % nd) % x) for x in v])
        got = tables.round_decimals(v, nd)
        np.testing.assert_array_equal(got, want)
        assert np.signbit(got).tolist() == np.signbit(want).tolist()
    rng = np.random.default_rng(3)
    v = rng.normal(size=20000) * 10
    np.testing.assert_array_equal(tables.round_decimals(v, 3), np.array([float("%.3f" % x) for x in v]))
    v32 = rng.normal(size=20000).astype(np.float32)
    np.testing.assert_array_equal(tables.round_decimals(v32, 5), np.array([float("%.5f" % x) for x in v32]))


@pytest.mark.parametrize("sep", [(1, 90), (1, 12), (12, 24), (24, 90), (3, 24)])
def test_select_equals_oracle(golden_dir, sep):
    params = tables.load_params()
    npz = np.load(f"{golden_dir}/example_NMR.npz")
    rst = tables.gen_rst(npz, params)
    seq = open(f"{golden_dir}/example_seq.fasta").read().split("\n")[1]
    for pcut, nogly in ((0.05, False), (0.15, True), (0.3, True)):
        params["PCUT"] = pcut
        got = tables.select(rst, sep[0], sep[1], params, seq, nogly)
        want = select_oracle(gen_rst_oracle(npz), sep[0], sep[1], pcut, seq, nogly)
        for name in got:
            np.testing.assert_array_equal(got[name], want[name])


def test_spline_knots_rules(golden_dir):
    params = tables.load_params()
    rst = tables.gen_rst(np.load(f"{golden_dir}/example_NMR.npz"), params)
    for name in tables.TYPES:
        for rule in ("H1", "H2"):
            x, y = tables.spline_knots(rst[name]["x"], rst[name]["y"], rst[name]["bin_size"], rule)
            xo, yo = apply_end_rule(rst[name]["x"], rst[name]["y"], rst[name]["bin_size"], rule)
            np.testing.assert_array_equal(x, xo)
            np.testing.assert_array_equal(y, yo)
    act = tables.active_restraints(rst, tables.select(rst, 1, 90, params))
    assert [len(act[t][0]) for t in tables.TYPES] == [3226, 2562, 5142, 2541]
    assert [len(act[t][2]) for t in tables.TYPES] == [37, 30, 30, 18]


# ---- restraint variants: product == oracle (the oracle is pinned on the reference's bytes in test_oracle_tables.py)

def _variant_inputs(golden_dir):
    g = np.load(f"{golden_dir}/gen_rst_variants24.npz")
    r = np.load(f"{golden_dir}/gen_rst_random24.npz")
    inp = {k: r[f"in_{k}"] for k in tables.TYPES}
    inp["idr"] = g["in_idr"]
    known = {k: g[f"in_known_{k}"] for k in ("dist", "omega", "theta_asym", "phi_asym")}
    return inp, known, dict(dist=g["in_af2_dist"], bins=g["in_af2_bins"])


def _same(got, want):
    assert list(got) == list(want)
    for name in want:
        for key in ("a", "b", "p", "x", "y"):
            np.testing.assert_array_equal(got[name][key], want[name][key], err_msg=f"{name} {key}")
        assert got[name]["bin_size"] == want[name]["bin_size"]


@pytest.mark.parametrize("orient", [True, False])
def test_variants_equal_oracle(golden_dir, orient):
    from oracle.tables_oracle import gen_idp_rst_oracle, gen_gpcr_rst_oracle
    params = tables.load_params()
    inp, known, af2 = _variant_inputs(golden_dir)
    _same(tables.gen_rst(inp, params, orient, "idp"), gen_idp_rst_oracle(inp, orient))
    _same(tables.gen_rst(inp, params, orient, "gpcr", known), gen_gpcr_rst_oracle(inp, known, orient))


def test_af2_variant_equals_oracle_and_flags_ca(golden_dir):
    from oracle.tables_oracle import gen_rst_af2_oracle
    params = tables.load_params()
    inp, known, af2 = _variant_inputs(golden_dir)
    got = tables.gen_rst(af2, params, False, "af2")
    _same(got, gen_rst_af2_oracle(af2))
    act = tables.active_restraints(got, None)
    assert act["dist_atom"] == "CA" and len(act["dist"][2]) == 62          # 60 listed knots + the two end knots of rule H1
    with pytest.raises(RuntimeError):                                       # the reference refuses orientations here (utils_ros.py:150)
        tables.gen_rst(af2, params, True, "af2")
    with pytest.raises(ValueError):
        tables.gen_rst(inp, params, True, "gpcr")                           # no -KNOWN


def test_select_idr_equals_oracle(golden_dir):
    from oracle.tables_oracle import gen_idp_rst_oracle, select_idr_oracle
    params = tables.load_params()
    inp, known, af2 = _variant_inputs(golden_dir)
    rst = tables.gen_rst(inp, params, True, "idp")
    seq = "AG" * 12
    for idr in (inp["idr"], 1 - inp["idr"]):
        for pcut, nogly in ((0.05, False), (0.15, True)):
            params["PCUT"] = pcut
            got = tables.select_idr(rst, idr, params, seq, nogly)
            want = select_idr_oracle(gen_idp_rst_oracle(inp, True), idr, pcut, seq, nogly)
            for name in got:
                np.testing.assert_array_equal(got[name], want[name])


# ---- edge cases of the table builder (empty / degenerate inputs)

def test_no_contacts_gives_empty_records_and_empty_selection():
    params = tables.load_params()
    L = 12
    def only_bin0(nb):
        a = np.zeros((L, L, nb), dtype=np.float32)
        a[..., 0] = 1.0                                     # every pair 'no contact': probability mass in bin 0
        return a
    npz = dict(dist=only_bin0(37), omega=only_bin0(25), theta=only_bin0(25), phi=only_bin0(13))
    rst = tables.gen_rst(npz, params)
    want = gen_rst_oracle(npz)
    for name in tables.TYPES:
        assert len(rst[name]["a"]) == 0 and rst[name]["y"].shape == (0, len(rst[name]["x"]))
        assert len(want[name]["a"]) == 0
    masks = tables.select(rst, 1, L, params)
    assert all(m.shape == (0,) for m in masks.values())
    act = tables.active_restraints(rst, masks)
    assert all(len(act[t][0]) == 0 and act[t][3].shape[0] == 0 for t in tables.TYPES)


def test_selection_window_edges_and_probability_thresholds(golden_dir):
    params = tables.load_params()
    rst = tables.gen_rst(np.load(f"{golden_dir}/example_NMR.npz"), params)
    # windows are half-open [sep1, sep2): a window of width zero selects nothing, adjacent windows partition
    assert not any(m.any() for m in tables.select(rst, 5, 5, params).values())
    a, b, c = (tables.select(rst, s1, s2, params) for s1, s2 in ((1, 12), (12, 24), (24, 90)))
    full = tables.select(rst, 1, 90, params)
    for name in tables.TYPES:
        assert not (a[name] & b[name]).any() and not (b[name] & c[name]).any()
        np.testing.assert_array_equal(a[name] | b[name] | c[name], full[name])
    # -pd acts on the selection only (gen_rst hard-codes 0.05, utils_ros.py:18): raising it can only remove restraints
    params["PCUT"] = 0.45
    high = tables.select(rst, 1, 90, params)
    for name in tables.TYPES:
        assert not (high[name] & ~full[name]).any() and high[name].sum() < full[name].sum()
    params["PCUT"] = 2.0
    assert not any(m.any() for m in tables.select(rst, 1, 90, params).values())


def test_idr_mask_extremes(golden_dir):
    from oracle.tables_oracle import gen_idp_rst_oracle
    params = tables.load_params()
    inp, known, af2 = _variant_inputs(golden_dir)
    base = tables.gen_rst(inp, params, True, "no-idp")
    none = dict(inp, idr=np.zeros_like(inp["idr"]))
    every = dict(inp, idr=np.ones_like(inp["idr"]))
    _same(tables.gen_rst(none, params, True, "idp"), base)                       # nothing flagged: gen_idp_rst == gen_rst
    _same(tables.gen_rst(every, params, True, "idp"), gen_idp_rst_oracle(every, True))
    # mode 3 with everything ordered: the 'disorder' stage adds nothing, the 'order' stage is add_rst over all separations
    rst = tables.gen_rst(none, params, True, "idp")
    first, second = tables.select_idr(rst, 1 - none["idr"], params), tables.select_idr(rst, none["idr"], params)
    everything = tables.select(rst, 0, 10 ** 6, params)
    for name in tables.TYPES:
        assert not second[name].any()
        np.testing.assert_array_equal(first[name], everything[name])
