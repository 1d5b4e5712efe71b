#!/usr/bin/env python
"""Drop-in for the reference's folding/folding.py (same command line, one PDB at -OUT).

run_inference.py -> utils_trX2dy/utils.py:484-505 shells out to
    python ./folding/folding.py -NPZ x.npz -FASTA x.fasta -OUT y.pdb -m 2 --orient -r no-idp
once per decoy.  This script accepts that exact command; the work (restraint tables,
NeRF, restraint + centroid energies, L-BFGS through the staged schedule) runs on the GPU
through libtrx2dyn.so.  With --ndecoy N the N independent decoys the reference would
produce with N processes (folding_with_pred_npz(repeat=N)) come out of ONE launch
(-OUT then holds a '{i}' placeholder).  There is no CPU fallback.

Not built: the Cartesian min_mover_cart stage and the full-atom FastRelax
stage (--fastrelax is accepted and ignored; decoys are centroid backbone + CB), and the
idp / af2 / gpcr restraint variants."""
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
    if args.rst != "no-idp":
        raise SystemExit("folding.py: restraint variant '-r %s' is not built in this version (only no-idp)" % args.rst)
    npz = np.load(args.NPZ)
    seq = read_fasta(args.FASTA)
    L = len(seq)
    params["seq"] = seq
    ctx = capi.Context(max(args.gpu, 0))
    rst = tables.gen_rst(npz, params)
    for name in ("dist", "omega", "theta", "phi"):
        if name in rst:
            print("%-6s restraints: %d" % (name, len(rst[name]["a"])))

    # separation windows of the modes (folding.py:125-171); restraints accumulate
    # (ConstraintSetMover.add_constraints(True)), so window k scores [first sep1, sep2_k)
    if args.mode == 0:
        windows = [(1, 12), (1, 24), (1, L)]
    elif args.mode == 1:
        windows = [(3, 24), (3, L)]
    elif args.mode == 2:
        windows = [(1, L)]
    else:
        raise SystemExit("folding.py: mode 3 needs the 'idr' order/disorder split, not built in this version")

    n = args.ndecoy
    seed = args.seed if args.seed is not None else int.from_bytes(os.urandom(4), "little")
    tors = sampler.random_torsions(n, L, seed)
    aa = sampler.aa_index(seq)  # Gly -> Ala for the centroid stage (folding.py:112-115)
    out = None
    for k, (s1, s2) in enumerate(windows):
        masks = tables.select(rst, s1, s2, params)
        tb = capi.Tables(ctx, L, tables.active_restraints(rst, masks, args.spline_end_rule))
        runs = schedule.reference_schedule() if k == 0 and len(windows) == 1 else schedule.window_schedule(initial_clash=(k == 0))
        batch = capi.FoldBatch(ctx, [tb], [n], aa, runs)
        out = batch.run(tors)
        tors = out["tors"]
        batch.close()
        tb.close()
    names = ("atom_pair_constraint", "dihedral_constraint", "angle_constraint", "vdw", "rama", "omega")
    for i in range(n):
        path = args.OUT.replace("{i}", str(args.start_id + i)) if n > 1 or "{i}" in args.OUT else args.OUT
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        remark = ["%s %.3f" % (nm, v) for nm, v in zip(names, out["terms"][i])]
        pdbio.write_pdb(path, seq, out["xyz"][i], remark)
    print("\ndone: %d decoy(s), %d energy evaluations each on average" % (n, int(out["evals"].mean())))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("*** time:%.2fs ***" % (time.time() - t0))
