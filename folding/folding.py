#!/usr/bin/env python
"""Drop-in for the reference's folding/folding.py (same command line, one PDB at -OUT).

run_inference.py -> utils_trX2dy/utils.py:484-505 shells out to
    python ./folding/folding.py -NPZ x.npz -FASTA x.fasta -OUT y.pdb -m 2 --orient -r no-idp
once per decoy.  This script accepts that exact command; the work (restraint tables,
NeRF, restraint + centroid energies, L-BFGS through the staged schedule) runs on the GPU
through libtrx2dyn.so.  With --ndecoy N the N independent decoys the reference would
produce with N processes (folding_with_pred_npz(repeat=N)) come out of ONE launch
(-OUT then holds a '{i}' placeholder).  There is no CPU fallback.

The staged schedule includes the Cartesian min_mover_cart stage.  Not built: the
full-atom FastRelax stage (--fastrelax is accepted and ignored; decoys are centroid backbone + CB) and
Rosetta's own database-driven potentials (vdw, rama, omega, cart_bonded, cen_hb / hbond are stated approximations)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from utils_ros.arguments import get_args  # noqa: E402


def read_fasta(path):
    """First chain of a FASTA file (folding.py:17-29)."""
    seq = ""
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if seq:
                    print("warning: multiple chains submitted, only the first is used", file=sys.stderr)
                    break
                continue
            seq += line.rstrip()
    return seq


def main(argv=None):
    import trx2dyn  # noqa: F401
    from trx2dyn import capi, pdbio, sampler, schedule, tables

    params = tables.load_params()
    args = get_args(params, argv)
    print(args)
    npz = np.load(args.NPZ)
    seq = read_fasta(args.FASTA)
    L = len(seq)
    params["seq"] = seq
    odd = sorted(set(seq) - set(sampler.AA_ORDER))
    if odd:
        print("warning: non-standard residue letter(s) %s are scored as Ala" % ", ".join(odd), file=sys.stderr)
    n = args.ndecoy
    if n > 1 and "{i}" not in args.OUT:
        raise SystemExit("folding.py: --ndecoy %d needs a '{i}' placeholder in -OUT (every decoy would be written to %s)" % (n, args.OUT))
    # what differs from the reference's energy function, said where a user sees it (and in the PDB REMARKs)
    dropped = schedule.ignored_terms()
    model_note = ["energy function: atom_pair/dihedral/angle constraints as the reference; vdw, rama, omega, cart_bonded and the",
                  "backbone H-bond term (for cen_hb / hbond_sr_bb / hbond_lr_bb) are stated approximations of Rosetta's",
                  "database-driven terms (include/trx_centroid_model.h); no full-atom FastRelax"]
    if dropped:
        model_note.append("weight-file terms NOT scored: " + ", ".join("%s %g (%s)" % (t, w, f) for f, t, w in dropped))
    for ln in model_note:
        print("note: " + ln, file=sys.stderr)
    ctx = capi.Context(max(args.gpu, 0))
    # restraint tables of the requested variant (folding.py:60-68)
    known = None
    if args.rst == "gpcr":
        if not args.KNOWN:
            raise SystemExit("folding.py: -r gpcr needs -KNOWN (npz of template maps: dist, omega, theta_asym, phi_asym)")
        known = np.load(args.KNOWN)
    rst = tables.gen_rst(npz, params, variant=args.rst, known=known)
    for name in ("dist", "omega", "theta", "phi"):
        if name in rst:
            print("%-6s restraints: %d" % (name, len(rst[name]["a"])))

    # restraint sets of the stages (folding.py:125-186); restraints accumulate
    # (ConstraintSetMover.add_constraints(True)), so stage k scores the union of stages <= k
    if args.mode == 0:
        stages = [tables.select(rst, 1, s2, params) for s2 in (12, 24, L)]
    elif args.mode == 1:
        stages = [tables.select(rst, 3, s2, params) for s2 in (24, L)]
    elif args.mode == 2:
        stages = [tables.select(rst, 1, L, params)]
    else:   # mode 3: ordered pairs first (odr = 1 - idr), then the disordered ones on top (folding.py:173-186)
        if "idr" not in npz:
            raise SystemExit("folding.py: -m 3 needs the 'idr' order/disorder map in the npz")
        idr = np.asarray(npz["idr"])
        first = tables.select_idr(rst, 1 - idr, params)
        second = tables.select_idr(rst, idr, params)
        stages = [first, {k: first[k] | second[k] for k in first}]

    seed = args.seed if args.seed is not None else int.from_bytes(os.urandom(4), "little")
    tors = sampler.random_torsions(n, L, seed)
    aa = sampler.aa_index(seq)  # Gly -> Ala for the centroid stage (folding.py:112-115)
    out = None
    for k, masks in enumerate(stages):
        # a stage with nothing selected still runs its movers (add_rst returns early, utils_ros.py:725, but
        # repeat_mover / min_mover_cart / remove_clash follow, folding.py:129-171): empty tables
        tb = capi.Tables(ctx, L, tables.active_restraints(rst, masks, args.spline_end_rule))
        # remove_clash(sf_vdw, min_mover_vdw) runs once, before the first stage (folding.py:119)
        runs = schedule.reference_schedule() if len(stages) == 1 else schedule.window_schedule(initial_clash=(out is None))
        batch = capi.FoldBatch(ctx, [tb], [n], aa, runs)
        out = batch.run(tors)
        tors = out["tors"]
        batch.close()
        tb.close()
    names = schedule.TERMS
    for i in range(n):
        path = args.OUT.replace("{i}", str(args.start_id + i)) if n > 1 or "{i}" in args.OUT else args.OUT
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        remark = ["%s %.3f" % (nm, v) for nm, v in zip(names, out["terms"][i])] + model_note
        pdbio.write_pdb(path, seq, out["xyz"][i], remark)
    print("\ndone: %d decoy(s), %d energy evaluations each on average" % (n, int(out["evals"].mean())))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("*** time:%.2fs ***" % (time.time() - t0))
