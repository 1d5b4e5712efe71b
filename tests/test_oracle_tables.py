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
