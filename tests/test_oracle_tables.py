"""Pins oracle/tables_oracle.py against the reference's own gen_rst output
(fixtures made by tests/golden/make_golden.py from /root/reference)."""
import hashlib
import numpy as np
import pytest

from oracle.tables_oracle import gen_rst_oracle, text_lines, select_oracle

TYPES = ("dist", "omega", "theta", "phi")


def _check(gold, rst, types):
    for name in types:
        rec = rst[name]
        np.testing.assert_array_equal(gold[f"{name}_a"], rec["a"])
        np.testing.assert_array_equal(gold[f"{name}_b"], rec["b"])
        np.testing.assert_array_equal(gold[f"{name}_p"], rec["p"])  # bit-exact float32
        h = hashlib.sha256()
        for k in range(len(rec["a"])):
            h.update("".join(text_lines(rec, k)).encode())
        assert h.hexdigest() == str(gold[f"{name}_sha256"]), name
        for k, t in zip(gold[f"{name}_sub_idx"], gold[f"{name}_sub_text"]):
            assert "".join(text_lines(rec, int(k))) == str(t)


def test_example_nmr_bytes_equal(golden_dir):
    gold = np.load(f"{golden_dir}/gen_rst_example_NMR.npz")
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    _check(gold, rst, TYPES)
    assert [len(rst[t]["a"]) for t in TYPES] == [3226, 3201, 6455, 4773]


def test_random24_bytes_equal(golden_dir):
    gold = np.load(f"{golden_dir}/gen_rst_random24.npz")
    inp = {k: gold[f"in_{k}"] for k in TYPES}
    _check(gold, gen_rst_oracle(inp), TYPES)
    gold2 = np.load(f"{golden_dir}/gen_rst_random24_noorient.npz")
    rst = gen_rst_oracle(inp, use_orient=False)
    assert list(rst) == ["dist"]
    _check(gold2, rst, ("dist",))


def test_selection_counts_example(golden_dir):
    # active after add_rst thresholds on the example NMR npz
    rst = gen_rst_oracle(np.load(f"{golden_dir}/example_NMR.npz"))
    sel = select_oracle(rst, 1, 90, 0.05)
    assert [int(sel[t].sum()) for t in TYPES] == [3226, 2562, 5142, 2541]
    sel = select_oracle(rst, 12, 24, 0.05)
    for t in TYPES:
        sep = np.abs(rst[t]["a"] - rst[t]["b"])[sel[t]]
        assert sep.min() >= 12 and sep.max() < 24


# ---- restraint variants: -r idp / af2 / gpcr against the reference's own output

def _variant_inputs(golden_dir):
    g = np.load(f"{golden_dir}/gen_rst_variants24.npz")
    r = np.load(f"{golden_dir}/gen_rst_random24.npz")
    inp = {k: r[f"in_{k}"] for k in TYPES}
    inp["idr"] = g["in_idr"]
    known = {k: g[f"in_known_{k}"] for k in ("dist", "omega", "theta_asym", "phi_asym")}
    af2 = dict(dist=g["in_af2_dist"], bins=g["in_af2_bins"])
    return g, inp, known, af2


def _check_variant(gold, tag, rst, tie_rows_ok=False):
    for name, rec in rst.items():
        pre = f"{tag}__{name}"
        np.testing.assert_array_equal(gold[f"{pre}_a"], rec["a"])
        np.testing.assert_array_equal(gold[f"{pre}_b"], rec["b"])
        np.testing.assert_array_equal(gold[f"{pre}_p"], rec["p"])
        texts = gold[f"{pre}_sub_text"]
        assert len(texts) == len(rec["a"])
        bad = [k for k in range(len(rec["a"])) if "".join(text_lines(rec, k)) != str(texts[k])]
        if not tie_rows_ok:
            assert not bad, (tag, name, bad[:5])
        else:
            # ling_sumlt picks the 5 lowest template knots with numpy's default (unstable) argsort; the padded
            # angular tables repeat knots, so which of two tied knots is taken depends on the numpy build / CPU.
            # The fixture was made in this container; elsewhere a few tied rows may legitimately differ.
            assert len(bad) <= 0.1 * len(rec["a"]), (tag, name, len(bad))
        assert str(gold[f"{pre}_line0"]).split()[0] == ("AtomPair" if name == "dist" else "Angle" if name == "phi" else "Dihedral")


def test_idp_variant_bytes_equal(golden_dir):
    from oracle.tables_oracle import gen_idp_rst_oracle
    gold, inp, known, af2 = _variant_inputs(golden_dir)
    _check_variant(gold, "idp", gen_idp_rst_oracle(inp, True))
    rst = gen_idp_rst_oracle(inp, False)
    assert list(rst) == ["dist"]
    _check_variant(gold, "idp_noorient", rst)
    # the variant really differs from gen_rst on flagged pairs and only there
    base = gen_rst_oracle(inp)
    for name in TYPES:
        flagged = inp["idr"][base[name]["a"], base[name]["b"]]
        diff = np.any(base[name]["y"] != gen_idp_rst_oracle(inp, True)[name]["y"], axis=1)
        assert diff[flagged].any() and not diff[~flagged].any()


def test_af2_variant_bytes_equal(golden_dir):
    from oracle.tables_oracle import gen_rst_af2_oracle
    gold, inp, known, af2 = _variant_inputs(golden_dir)
    rst = gen_rst_af2_oracle(af2)
    _check_variant(gold, "af2", rst)
    assert rst["dist"]["atom"] == "CA" and len(rst["dist"]["x"]) == 60 and rst["dist"]["bin_size"] == 0.3125
    assert str(gold["af2__dist_line0"]).startswith("AtomPair CA 1 CA ")


def test_gpcr_variant_bytes_equal(golden_dir):
    from oracle.tables_oracle import gen_gpcr_rst_oracle
    gold, inp, known, af2 = _variant_inputs(golden_dir)
    _check_variant(gold, "gpcr", gen_gpcr_rst_oracle(inp, known, True), tie_rows_ok=True)
    _check_variant(gold, "gpcr_noorient", gen_gpcr_rst_oracle(inp, known, False), tie_rows_ok=True)


def test_mode3_selection(golden_dir):
    from oracle.tables_oracle import gen_idp_rst_oracle, select_idr_oracle
    gold, inp, known, af2 = _variant_inputs(golden_dir)
    rst = gen_idp_rst_oracle(inp, True)
    idr = inp["idr"]
    dis, ordr = select_idr_oracle(rst, idr, 0.05), select_idr_oracle(rst, 1 - idr, 0.05)
    every = select_oracle(rst, 0, 10 ** 6, 0.05)
    for name in TYPES:
        assert not (dis[name] & ordr[name]).any()
        np.testing.assert_array_equal(dis[name] | ordr[name], every[name])   # the two stages of mode 3 partition add_rst's set
        assert dis[name].any() and ordr[name].any()
