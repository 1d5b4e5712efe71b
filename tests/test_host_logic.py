"""Host-side logic that needs no GPU: schedule from the .wts files, random starts,
PDB writer contract, CLI parsing."""
import os
import sys

import numpy as np
import pytest

import trx2dyn  # noqa: F401
from trx2dyn import pdbio, sampler, schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_matches_reference_weights_and_caps():
    runs = schedule.reference_schedule()
    # remove_clash (<=5 x, max_iter 500, rama+vdw) ; 3 x min_mover (1000) ; min_mover_cart (1000, Cartesian) ;
    # remove_clash with sf1 (<=5 x, 1000)            folding.py:86-104,118-119,164-171
    assert len(runs) == 14
    assert [r.max_iter for r in runs] == [500] * 5 + [1000] * 9
    # last entry: the backbone hydrogen-bond term that stands in for cen_hb / hbond_sr_bb / hbond_lr_bb
    assert list(runs[0].w) == [0, 0, 0, 1, 1, 0, 0, 0]              # scorefxn_vdw.wts
    assert list(runs[5].w) == [5, 4, 4, 1, 1, 0.5, 0, 5]            # scorefxn.wts (cen_hb 5)
    assert list(runs[8].w) == [5, 4, 4, 0.5, 1, 0.5, 0.1, 3]        # scorefxn_cart.wts (hbond_sr_bb 3, hbond_lr_bb 3)
    assert list(runs[9].w) == [3, 1, 1, 3, 1, 0.5, 0, 5]            # scorefxn1.wts (cen_hb 5)
    assert schedule.ignored_terms() == []                           # every weight-file term is scored (as a stated approximation or exactly)
    assert all(r.tol == 1e-4 for r in runs)
    assert [r.cartesian for r in runs] == [0] * 8 + [1] + [0] * 5
    assert [r.clash_check for r in runs] == [1] * 5 + [0] * 4 + [1] * 5
    assert all(r.skip_to == 5 for r in runs[:5]) and all(r.skip_to == 14 for r in runs[9:])
    assert all(r.clash_thr == 10.0 for r in runs)
    # the torsion-only variant and the per-window block of modes 0/1
    assert [r.cartesian for r in schedule.reference_schedule(cartesian=False)] == [0] * 13
    win = schedule.window_schedule(initial_clash=False)
    assert [r.cartesian for r in win] == [0, 0, 0, 1] + [0] * 5 and all(r.skip_to == 9 for r in win[4:])


def test_random_torsions_follow_the_six_state_table():
    t = np.rad2deg(sampler.random_torsions(4000, 30, seed=1).astype(np.float64))
    assert np.allclose(t[:, :, 2], 180.0) and np.allclose(t[:, 29, :], 180.0)
    states = {(-140, 153): 0.135, (-72, 145): 0.155, (-122, 117): 0.073, (-82, -14): 0.122, (-61, -41): 0.497, (57, 39): 0.018}
    pairs = np.round(t[:, :29, :2]).reshape(-1, 2)
    for (phi, psi), p in states.items():
        frac = np.mean((pairs[:, 0] == phi) & (pairs[:, 1] == psi))
        assert abs(frac - p) < 0.01
    a = sampler.random_torsions(3, 10, seed=5)
    np.testing.assert_array_equal(a, sampler.random_torsions(3, 10, seed=5))
    assert sampler.aa_index("AGP").tolist() == [0, 0, 14]         # Gly scored as Ala


def test_pdb_writer_contract(tmp_path):
    seq = "MGAKW"
    xyz = np.random.default_rng(0).normal(size=(5, 5, 3)) * 10
    p = tmp_path / "d.pdb"
    pdbio.write_pdb(str(p), seq, xyz, ["vdw 1.0"])
    text = p.read_text()
    atoms = [ln for ln in text.splitlines() if ln.startswith("ATOM")]
    assert len(atoms) == 5 * 5 - 1                                  # Gly has no CB
    assert all(len(ln) == 80 for ln in atoms)
    assert [ln[12:16] for ln in atoms[:5]] == [" N  ", " CA ", " C  ", " O  ", " CB "]
    seq2, at = pdbio.read_backbone(str(p))
    assert seq2 == seq
    assert np.isnan(at["CB"][1]).all() and not np.isnan(at["CB"][0]).any()
    np.testing.assert_allclose(at["CA"], np.round(xyz[:, 1], 3), atol=1e-9)
    np.testing.assert_allclose(at["C"], np.round(xyz[:, 3], 3), atol=1e-9)
    assert [int(ln[22:26]) for ln in atoms if ln[12:16] == " CA "] == [1, 2, 3, 4, 5]


def test_cli_accepts_the_reference_command_line():
    sys.path.insert(0, os.path.join(ROOT, "folding"))
    from utils_ros.arguments import get_args
    params = {"PCUT": 0.05, "WDIR": "/dev/shm"}
    # utils_trX2dy/utils.py:491-498 + run_inference.py:295
    a = get_args(params, ["-NPZ", "x.npz", "-FASTA", "x.fasta", "-OUT", "o.pdb", "-m", "2", "--orient", "-r", "no-idp"])
    assert (a.mode, a.rst, a.use_orient, a.fastrelax, a.pcut) == (2, "no-idp", True, True, 0.05)
    assert params["USE_ORIENT"] is True
    a = get_args(params, ["-NPZ", "x", "-FASTA", "y", "-OUT", "z", "-m", "2", "--no-orient", "-r", "no-idp", "-pd", "0.15", "--no-fastrelax"])
    assert (a.use_orient, a.fastrelax, params["PCUT"]) == (False, False, 0.15)
