"""In-process equivalent of run_inference.run_single AFTER the network has written its npz
files (run_inference.py:298-336): for each model's distograms, fold init_num decoys in one
GPU batch, keep the most reliable one, then iterate decay -> fold until the distogram stops
changing or Nmax decoys, and write the decoys as save_dir/name/pred_pdb/conf_{model}_{k}.pdb.

Naming: the reference flattens NMR/ and Xray/ into one directory and renames initial{i}.pdb ->
conf_1_{i+1}, initial{i}_1.pdb -> conf_2_{i+1}, and the iteration decoys after them
(run_inference.py:145-278; its lexicographic sort mis-attributes iteration decoys beyond the
9th -- here every decoy keeps the model it came from)."""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import capi, dynamics, pdbio, sampler, tables


def read_fasta(path):
    seq = ""
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if seq:
                    break
                continue
            seq += line.strip()
    return seq


def run_single_from_npz(name, fasta, npz_paths, save_dir, init_num=10, n_max=300, angle=True, device=0, seed=0,
                        rule="H1", ctx=None):
    """npz_paths: [NMR npz] or [NMR npz, X-ray npz] (--mult_two_models).  Returns the list of PDB paths."""
    seq = read_fasta(fasta)
    L = len(seq)
    own = ctx is None
    ctx = ctx or capi.Context(device)
    out_dir = os.path.join(save_dir, name, "pred_pdb")
    os.makedirs(out_dir, exist_ok=True)
    params = tables.load_params()
    params["USE_ORIENT"] = bool(angle)
    written = []
    for m, path in enumerate(npz_paths, start=1):
        npz0 = {k: np.asarray(v) for k, v in np.load(path).items()}
        counter = {"k": 0, "calls": 0}

        def fold_fn(npz, n):
            counter["calls"] += 1
            return sampler.fold(ctx, [npz], seq, [n], seed=seed + 7919 * m + counter["calls"], params=params, rule=rule)

        def on_decoy(tag, xyz):
            counter["k"] += 1
            p = os.path.join(out_dir, "conf_%d_%d.pdb" % (m, counter["k"]))
            pdbio.write_pdb(p, seq, xyz, ["source %s model %d" % (tag, m)])
            written.append(p)

        dynamics.generate(fold_fn, npz0, L, n_init=init_num, n_max=n_max, angle=angle, on_decoy=on_decoy)
    if own:
        ctx.close()
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description="fold + dynamics loop from predicted npz files (GPU)")
    ap.add_argument("--fasta", required=True)
    ap.add_argument("--npz", nargs="+", required=True, help="predicted distograms: NMR [X-ray]")
    ap.add_argument("--name", default="seq")
    ap.add_argument("--save_dir", default="./output")
    ap.add_argument("--init_num", type=int, default=10)
    ap.add_argument("--Nmax", type=int, default=300)
    ap.add_argument("--angle", dest="angle", action="store_true", default=True)
    ap.add_argument("--no-angle", dest="angle", action="store_false")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    files = run_single_from_npz(a.name, a.fasta, a.npz, a.save_dir, a.init_num, a.Nmax, a.angle, a.device, a.seed)
    print("wrote %d decoys under %s" % (len(files), os.path.join(a.save_dir, a.name, "pred_pdb")))


if __name__ == "__main__":
    main()
