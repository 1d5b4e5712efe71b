"""The reference's minimisation schedule as data for the device state machine.

folding/folding.py:74-104 builds four ScoreFunctions from data/*.wts and four MinMovers
('lbfgs_armijo_nonmonotone', tolerance 1e-4, max_iter 1000/1000/500/1000);
:118-119 and :164-171 (mode 2) apply them:
    remove_clash(sf_vdw, min_mover_vdw)        <= 5 x { if sf_vdw(pose) < 10: break; minimise }
    RepeatMover(min_mover, 3)                  3 x minimise under scorefxn.wts
    min_mover_cart                             Cartesian stage: xyz are the degrees of freedom (scorefxn_cart.wts)
    remove_clash(sf_vdw, min_mover1)           <= 5 x { ...; minimise under scorefxn1.wts }
Terms the library implements: atom_pair_constraint, dihedral_constraint, angle_constraint,
vdw, rama, omega, cart_bonded and one backbone hydrogen-bond term for cen_hb / hbond_sr_bb / hbond_lr_bb
(the non-restraint ones as stated approximations: Rosetta's are database-driven and not in the reference tree)."""
from __future__ import annotations

import os

from .capi import Run, NTERM

TERMS = ("atom_pair_constraint", "dihedral_constraint", "angle_constraint", "vdw", "rama", "omega", "cart_bonded", "backbone_hbond")
# Rosetta's backbone hydrogen-bond terms all map onto the one stated approximation of include/trx_centroid_model.h:
# cen_hb in the centroid stages; hbond_sr_bb / hbond_lr_bb (the same potential on complementary sequence
# separations, the same weight in scorefxn_cart.wts) in the Cartesian stage
_HB_ALIASES = ("cen_hb", "hbond_sr_bb", "hbond_lr_bb")
_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "folding", "data")


def read_wts(name, data_dir=None):
    """Weights file -> list of the NTERM weights this library scores (others ignored)."""
    w = dict.fromkeys(TERMS, 0.0)
    with open(os.path.join(data_dir or _DATA, name)) as fh:
        for line in fh:
            tok = line.split()
            if len(tok) >= 2 and tok[0] in w:
                w[tok[0]] = float(tok[1])
            elif len(tok) >= 2 and tok[0] in _HB_ALIASES:
                w["backbone_hbond"] = max(w["backbone_hbond"], float(tok[1]))
    return [w[t] for t in TERMS]


def ignored_terms(data_dir=None, names=("scorefxn.wts", "scorefxn1.wts", "scorefxn_cart.wts", "scorefxn_vdw.wts")):
    """[(file, term, weight)] of the weight-file entries this library does NOT score (Rosetta's database-driven
    cen_hb / hbond_sr_bb / hbond_lr_bb: folding/data/scorefxn.wts:1, scorefxn1.wts:1, scorefxn_cart.wts:1-2).
    The drop-in folding.py prints them and records them in the PDB REMARKs: its decoys come from an energy
    function that differs from the reference's in these terms, and whose vdw / rama / omega / cart_bonded are
    stated approximations (include/trx_centroid_model.h)."""
    out = []
    for name in names:
        with open(os.path.join(data_dir or _DATA, name)) as fh:
            for line in fh:
                tok = line.split()
                if len(tok) >= 2 and tok[0] not in TERMS and tok[0] not in _HB_ALIASES and not tok[0].startswith("#"):
                    try:
                        out.append((name, tok[0], float(tok[1])))
                    except ValueError:
                        pass
    return out


def make_run(w, max_iter, tol=1e-4, clash_check=False, clash_thr=10.0, skip_to=0, cartesian=False):
    r = Run()
    for k in range(NTERM):
        r.w[k] = w[k]
    r.max_iter, r.tol = int(max_iter), float(tol)
    r.clash_check, r.clash_thr, r.skip_to, r.cartesian = int(clash_check), float(clash_thr), int(skip_to), int(cartesian)
    return r


def _stage(runs, data_dir, cartesian):
    """{RepeatMover(min_mover,3); min_mover_cart; remove_clash(sf_vdw, min_mover1)} appended to runs
    (folding.py:129-171: the block every mode repeats after add_rst)."""
    sf = read_wts("scorefxn.wts", data_dir)
    sf1 = read_wts("scorefxn1.wts", data_dir)
    runs += [make_run(sf, 1000) for _ in range(3)]
    if cartesian:
        runs.append(make_run(read_wts("scorefxn_cart.wts", data_dir), 1000, cartesian=True))
    end = len(runs) + 5
    runs += [make_run(sf1, 1000, clash_check=True, skip_to=end) for _ in range(5)]
    return runs


def reference_schedule(data_dir=None, cartesian=True):
    """Mode-2 schedule.  Modes 0 and 1 of folding.py:125-160 repeat the {repeat_mover, cart,
    remove_clash} block per separation window; the host driver then calls the fold once per
    window with the window's tables (window_schedule)."""
    sf_vdw = read_wts("scorefxn_vdw.wts", data_dir)
    runs = [make_run(sf_vdw, 500, clash_check=True, skip_to=5) for _ in range(5)]
    return _stage(runs, data_dir, cartesian)


def window_schedule(data_dir=None, initial_clash=False, cartesian=True):
    """One separation window of modes 0/1/3: RepeatMover(min_mover,3) + min_mover_cart + remove_clash(min_mover1)."""
    sf_vdw = read_wts("scorefxn_vdw.wts", data_dir)
    runs = []
    if initial_clash:
        runs += [make_run(sf_vdw, 500, clash_check=True, skip_to=5) for _ in range(5)]
    return _stage(runs, data_dir, cartesian)


def mc_schedule(data_dir=None, mc_max_iter=200, cartesian=True):
    """reference_schedule() followed by the run Monte-Carlo cycles re-minimise with and score
    on (scorefxn.wts, shorter cap).  An extension: the reference has no MC step in folding/."""
    runs = reference_schedule(data_dir, cartesian)
    runs.append(make_run(read_wts("scorefxn.wts", data_dir), mc_max_iter))
    return runs
