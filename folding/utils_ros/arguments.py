"""Command line of folding.py: the reference's flags (folding/utils_ros/arguments.py:6-25,
same names, defaults and destinations) plus the batched extras of this build."""
import argparse


def get_args(params, argv=None):
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-NPZ", type=str, required=True, help="input distograms and anglegrams (NN predictions)")
    p.add_argument("-FASTA", type=str, required=True, help="input sequence")
    p.add_argument("-OUT", type=str, required=True, help="output model (in PDB format)")
    p.add_argument("-KNOWN", type=str, required=False, help="if r=gpcr input known pdb")
    p.add_argument("-pd", type=float, dest="pcut", default=params["PCUT"], help="min probability of distance restraints")
    p.add_argument("-m", type=int, dest="mode", default=2, choices=[0, 1, 2, 3], help="0: sh+m+l, 1: (sh+m)+l, 2: (sh+m+l)")
    p.add_argument("-r", type=str, dest="rst", default="no-idp", choices=["no-idp", "idp", "gpcr", "af2"],
                   help="add rst:no-idp:order,idp:disorder,gpcr:two conf,af2:af2 bins")
    p.add_argument("-w", type=str, dest="wdir", default=params["WDIR"], help="folder to store temp files (unused: nothing is written)")
    p.add_argument("-n", type=int, dest="steps", default=1000, help="number of minimization steps (unused, as in the reference)")
    p.add_argument("--orient", dest="use_orient", action="store_true", help="use orientations")
    p.add_argument("--no-orient", dest="use_orient", action="store_false")
    p.add_argument("--fastrelax", dest="fastrelax", action="store_true", help="accepted; the full-atom stage is out of scope, decoys stay centroid")
    p.add_argument("--no-fastrelax", dest="fastrelax", action="store_false")
    p.add_argument("--log", dest="log", default="")
    p.add_argument("--gpu", dest="gpu", default=-1, type=int, help="CUDA device (default: device 0)")
    # extras of this build
    p.add_argument("--ndecoy", type=int, default=1, help="fold this many decoys in one launch; -OUT may contain {i}")
    p.add_argument("--start-id", type=int, default=0, help="first value of {i}")
    p.add_argument("--seed", type=int, default=None, help="seed of the random phi/psi starts (reference: unseeded)")
    p.add_argument("--spline-end-rule", default="H1", choices=["H1", "H2"], help="SplineFunc end-knot rule (DESIGN.md)")
    p.set_defaults(use_orient=True)
    p.set_defaults(fastrelax=True)
    args = p.parse_args(argv)
    params["PCUT"] = args.pcut
    params["USE_ORIENT"] = args.use_orient
    return args
