"""PyRosetta pinning.  These tests ACTIVATE when tests/golden/pyrosetta_*.npz exist -- vectors made
on a host with PyRosetta by tools/pyrosetta_golden.py (PyRosetta is absent from this image and from the
reference tree).  They settle the SplineFunc end-knot rule (H1 vs H2) and hold the oracle and
the fp64 kernel to the north-star tolerances against PyRosetta itself: energies 1e-6 relative, gradients 1e-5
relative (of the largest component).  Without the vectors they skip, and parity below the tables stays
"unpinned against PyRosetta" (DESIGN.md)."""
import glob
import os

import numpy as np
import pytest

import trx2dyn  # noqa: F401
from oracle import restraints_oracle as ro
from oracle.tables_oracle import gen_rst_oracle, select_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "pyrosetta_*.npz")))
needs_vectors = pytest.mark.skipif(not FILES, reason="no tests/golden/pyrosetta_*.npz: run tools/pyrosetta_golden.py on a host with PyRosetta")


def _sets():
    npz = np.load(os.path.join(GOLD, "example_NMR.npz"))
    rst = gen_rst_oracle(npz)
    sel = select_oracle(rst, 1, npz["dist"].shape[0], 0.05)
    return rst, sel


def _rule(rst, sel):
    """The end-knot rule PyRosetta's SplineFunc follows: the one under which the oracle reproduces its energies."""
    err = {}
    for rule in ("H1", "H2"):
        rs = ro.RestraintSetOracle(rst, sel, rule)
        e = 0.0
        for f in FILES:
            g = np.load(f)
            E, _ = rs.energy_grad(g["xyz"], (1.0, 1.0, 1.0))
            e = max(e, float(np.max(np.abs(E - g["E"]) / np.maximum(np.abs(g["E"]), 1.0))))
        err[rule] = e
    return min(err, key=err.get), err


def test_kit_is_present_and_self_contained():
    # the kit itself is part of the repository whether or not its output is
    src = open(os.path.join(os.path.dirname(GOLD), "..", "tools", "pyrosetta_golden.py")).read()
    assert "atom_pair_constraint" in src and "dihedral_constraint" in src and "angle_constraint" in src
    assert "import pyrosetta" in src and "/root/reference" not in src.replace("/path/to/reference", "")


def test_kit_writes_the_references_restraint_files(tmp_path):
    """What the kit feeds PyRosetta is what the reference feeds it: constraint lines and spline files byte-identical to
    a run of the reference's own gen_rst on the example npz (hashes in tests/golden/gen_rst_example_NMR.npz)."""
    import hashlib
    import importlib.util
    spec = importlib.util.spec_from_file_location("pyrosetta_golden", os.path.join(os.path.dirname(GOLD), "..", "tools", "pyrosetta_golden.py"))
    kit = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kit)                      # importing the kit does not import pyrosetta (only main() does)
    g = np.load(os.path.join(GOLD, "gen_rst_example_NMR.npz"))
    rst, _ = _sets()
    every = {name: np.ones(len(rst[name]["a"]), dtype=bool) for name in rst}
    cst, n = kit.write_restraints(rst, every, str(tmp_path))
    lines = open(cst).read().splitlines()
    assert n == len(lines) == sum(len(rst[k]["a"]) for k in rst)
    pos = 0
    for name in ("dist", "omega", "theta", "phi"):
        cnt = len(rst[name]["a"])
        hl, ht = hashlib.sha256(), hashlib.sha256()
        for ln in lines[pos:pos + cnt]:
            hl.update(ln.replace(str(tmp_path), "TMP").encode())
            path = [tok for tok in ln.split() if tok.startswith(str(tmp_path))][0]
            ht.update(open(path).read().encode())
        assert hl.hexdigest() == str(g[f"{name}_lines_sha256"]), name
        assert ht.hexdigest() == str(g[f"{name}_sha256"]), name
        pos += cnt


@needs_vectors
def test_oracle_matches_pyrosetta_and_settles_the_end_knot_rule():
    rst, sel = _sets()
    rule, err = _rule(rst, sel)
    assert err[rule] < 1e-6, err                       # north_star: energies within 1e-6 relative
    assert err["H2" if rule == "H1" else "H1"] > 1e-5, err   # the other hypothesis is visibly off
    rs = ro.RestraintSetOracle(rst, sel, rule)
    for f in FILES:
        g = np.load(f)
        _, grad = rs.energy_grad(g["xyz"], (1.0, 1.0, 1.0))
        ref = g["grad"]
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max(), f     # gradients within 1e-5 relative
    # the library's default rule must be the one PyRosetta follows
    from trx2dyn import tables
    import inspect
    assert inspect.signature(tables.active_restraints).parameters["rule"].default == rule


@needs_vectors
@pytest.mark.gpu
def test_fp64_kernel_matches_pyrosetta():
    from trx2dyn import capi, tables
    rst_o, sel = _sets()
    rule, _ = _rule(rst_o, sel)
    npz = np.load(os.path.join(GOLD, "example_NMR.npz"))
    params = tables.load_params()
    rst = tables.gen_rst(npz, params)
    L = npz["dist"].shape[0]
    ctx = capi.Context(0)
    tb = capi.Tables(ctx, L, tables.active_restraints(rst, tables.select(rst, 1, L, params), rule))
    for f in FILES:
        g = np.load(f)
        E, grad = tb.energy_grad(g["xyz"][None], (1.0, 1.0, 1.0), capi.F64)
        assert np.max(np.abs(E[0] - g["E"]) / np.maximum(np.abs(g["E"]), 1.0)) < 1e-6, f
        assert np.abs(grad[0] - g["grad"]).max() <= 1e-5 * np.abs(g["grad"]).max(), f
    tb.close()
    ctx.close()
