#!/usr/bin/env python
"""PyRosetta pinning kit (SURVEY.md 8c "one-time external step"): run ONCE on any host that has PyRosetta.

    python tools/pyrosetta_golden.py [--pdb /path/to/reference/example/output/seq/pred_pdb/conf_1_1.pdb]
                                     [--out tests/golden]

It scores three coordinate sets of the reference's example target (L = 90, tests/golden/example_NMR.npz,
example_seq.fasta) with a ScoreFunction that holds ONLY atom_pair_constraint, dihedral_constraint and
angle_constraint at weight 1, using the restraint files exactly as the reference writes them
(folding/utils_ros/utils_ros.py:6-146: 'x_axis ...' / 'y_axis ...' text files, AtomPair / Dihedral / Angle
SPLINE lines; add_rst's selection, utils_ros.py:706-743), and dumps, per coordinate set,

    xyz   (L,3,3)  N, CA, CB as Rosetta holds them (centroid pose, Gly mutated to Ala as folding.py:112-115)
    E     (3,)     atom_pair_constraint, dihedral_constraint, angle_constraint (unweighted totals)
    grad  (L,3,3)  dE/dx of the three terms together, weights (1,1,1): central differences on the pose
                   (step 1e-4 A: fp64 Rosetta scores make this good to ~1e-7 relative), and, when this
                   PyRosetta build exposes it, the analytic per-atom F2 of the derivative pass ('grad_f2')

into tests/golden/pyrosetta_<set>.npz.  Sets: 'decoy' (--pdb: one of the reference's own example decoys;
skipped without --pdb), 'random' (the 6-state random phi/psi start of utils_ros.py:656-696, seed 0, built by
Rosetta from ideal geometry) and 'extended' (phi = psi = omega = 180).  With those files committed,
tests/test_pyrosetta_golden.py activates: it settles the SplineFunc end-knot rule (H1 vs H2, SURVEY 8a
row 9) and checks the oracle and the fp64 kernel against PyRosetta at 1e-6 (energies) / 1e-5 (gradients).

Nothing in the product path or the default test-suite needs PyRosetta; this script is the only place that
imports it."""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_SUFFIX = {"dist": "", "omega": "_omega", "theta": "_theta", "phi": "_phi"}


def constraint_line(name, a, b, path):
    """The line gen_rst builds for one record (utils_ros.py:73,95,114,138), 1-based residue numbers."""
    if name == "dist":
        return "AtomPair %s %d %s %d SPLINE TAG %s 1.0 %.3f %.5f" % ("CB", a, "CB", b, path, 1.0, 0.5)
    if name == "omega":
        return "Dihedral CA %d CB %d CB %d CA %d SPLINE TAG %s 1.0 %.3f %.5f" % (a, a, b, b, path, 1.0, np.deg2rad(15.0))
    if name == "theta":
        return "Dihedral N %d CA %d CB %d CB %d SPLINE TAG %s 1.0 %.3f %.5f" % (a, a, a, b, path, 1.0, np.deg2rad(15.0))
    return "Angle CA %d CB %d CB %d SPLINE TAG %s 1.0 %.3f %.5f" % (a, a, b, path, 1.0, np.deg2rad(15.0))


def write_restraints(rst, masks, tmpdir):
    """The reference's on-disk form of the selected restraints: one two-line spline file per restraint
    (utils_ros.py:62-72,88-94,108-113,132-137: 'x_axis ...' / 'y_axis ...', named a.b.txt, a.b_omega.txt, ...) and one
    constraint line each in minimize.cst (add_rst, utils_ros.py:731-735).  Files and lines are byte-identical to what
    the reference writes (tests/test_pyrosetta_golden.py checks both against hashes of a reference run)."""
    from oracle.tables_oracle import text_lines
    lines = []
    for name in ("dist", "omega", "theta", "phi"):
        if name not in rst:
            continue
        rec = rst[name]
        for k in np.nonzero(masks[name])[0]:
            a, b = int(rec["a"][k]) + 1, int(rec["b"][k]) + 1
            path = os.path.join(tmpdir, "%d.%d%s.txt" % (a, b, _SUFFIX[name]))
            with open(path, "w") as fh:
                fh.writelines(text_lines(rec, k))
            lines.append(constraint_line(name, a, b, path) + "\n")
    cst = os.path.join(tmpdir, "minimize.cst")
    with open(cst, "w") as fh:
        fh.writelines(lines)
    return cst, len(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdb", default=None, help="a reference example decoy, e.g. example/output/seq/pred_pdb/conf_1_1.pdb")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--fd-step", type=float, default=1e-4)
    args = ap.parse_args()

    import pyrosetta
    from pyrosetta import rosetta
    pyrosetta.init("-hb_cen_soft -relax:default_repeats 5 -default_max_cycles 200 -out:level 100 -mute all")   # folding.py:48

    from oracle.tables_oracle import gen_rst_oracle, select_oracle
    from trx2dyn import sampler  # noqa: E402  (only the seeded 6-state start)
    gold = os.path.join(ROOT, "tests", "golden")
    seq = open(os.path.join(gold, "example_seq.fasta")).read().split("\n")[1].strip()
    L = len(seq)
    npz = np.load(os.path.join(gold, "example_NMR.npz"))
    rst = gen_rst_oracle(npz)                       # byte-exact with the reference's gen_rst (tests/test_oracle_tables.py)
    masks = select_oracle(rst, 1, L, 0.05)          # mode 2: add_rst(pose, rst, 1, len(seq), params), folding.py:168

    sf = rosetta.core.scoring.ScoreFunction()
    st = rosetta.core.scoring.ScoreType
    for term in (st.atom_pair_constraint, st.dihedral_constraint, st.angle_constraint):
        sf.set_weight(term, 1.0)

    def centroid_pose_from_sequence():
        pose = pyrosetta.pose_from_sequence(seq, "centroid")          # folding.py:109
        return pose

    def mutate_gly(pose):
        for i, a in enumerate(seq):                                    # folding.py:112-115
            if a == "G":
                rosetta.protocols.simple_moves.MutateResidue(i + 1, "ALA").apply(pose)

    poses = {}
    p = centroid_pose_from_sequence()
    for i in range(1, L + 1):
        p.set_phi(i, 180.0); p.set_psi(i, 180.0); p.set_omega(i, 180.0)
    poses["extended"] = p
    p = centroid_pose_from_sequence()
    t = np.rad2deg(sampler.random_torsions(1, L, 0)[0])
    for i in range(1, L + 1):
        p.set_phi(i, float(t[i - 1, 0])); p.set_psi(i, float(t[i - 1, 1])); p.set_omega(i, float(t[i - 1, 2]))
    poses["random"] = p
    if args.pdb:
        p = pyrosetta.pose_from_pdb(args.pdb)
        rosetta.protocols.simple_moves.SwitchResidueTypeSetMover("centroid").apply(p)
        poses["decoy"] = p

    os.makedirs(args.out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        cst, n_cst = write_restraints(rst, masks, tmp)
        for tag, pose in poses.items():
            mutate_gly(pose)
            mover = rosetta.protocols.constraint_movers.ConstraintSetMover()          # utils_ros.py:738-741
            mover.constraint_file(cst)
            mover.add_constraints(True)
            mover.apply(pose)
            sf(pose)
            en = pose.energies().total_energies()
            E = np.array([en[st.atom_pair_constraint], en[st.dihedral_constraint], en[st.angle_constraint]])
            xyz = np.zeros((L, 3, 3))
            for i in range(L):
                r = pose.residue(i + 1)
                for k, name in enumerate(("N", "CA", "CB")):
                    v = r.xyz(name)
                    xyz[i, k] = (v.x, v.y, v.z)
            # central differences on the pose itself
            grad = np.zeros((L, 3, 3))
            h = args.fd_step
            V = rosetta.numeric.xyzVector_double_t
            for i in range(L):
                for k, name in enumerate(("N", "CA", "CB")):
                    aid = rosetta.core.id.AtomID(pose.residue(i + 1).atom_index(name), i + 1)
                    for c in range(3):
                        x0 = xyz[i, k].copy()
                        e = []
                        for sgn in (+1.0, -1.0):
                            x1 = x0.copy(); x1[c] += sgn * h
                            pose.set_xyz(aid, V(*x1))
                            e.append(sf(pose))
                        pose.set_xyz(aid, V(*x0))
                        grad[i, k, c] = (e[0] - e[1]) / (2 * h)
            out = dict(xyz=xyz, E=E, grad=grad, n_restraints=n_cst, fd_step=h, seq=seq,
                       pyrosetta_version=str(getattr(pyrosetta, "__version__", "unknown")))
            # analytic per-atom F2 of Rosetta's derivative pass, when this build exposes the call
            try:
                sf(pose)
                sf.setup_for_derivatives(pose)
                dm = pose.energies().domain_map()
                f2 = np.zeros((L, 3, 3))
                for i in range(L):
                    for k, name in enumerate(("N", "CA", "CB")):
                        aid = rosetta.core.id.AtomID(pose.residue(i + 1).atom_index(name), i + 1)
                        F1, F2 = V(0, 0, 0), V(0, 0, 0)
                        sf.eval_npd_atom_derivative(aid, pose, dm, F1, F2)
                        f2[i, k] = (F2.x, F2.y, F2.z)
                out["grad_f2"] = f2
            except Exception as exc:   # noqa: BLE001 -- API differs between PyRosetta builds; the FD gradient stands
                out["grad_f2_error"] = repr(exc)
            path = os.path.join(args.out, "pyrosetta_%s.npz" % tag)
            np.savez_compressed(path, **out)
            print("%s: %d restraints, E = %s -> %s" % (tag, n_cst, E, path))


if __name__ == "__main__":
    main()
