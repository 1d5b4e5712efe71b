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


def _run_chain(ctx, seq, npz_path, m, out_dir, init_num, n_max, angle, seed, rule, params):
    """One model's chain of run_inference.generate_npz_and_pdb (run_inference.py:16-143): init_num decoys in one
    batch, then one decoy per iteration on the decayed distograms.  Iterations depend on each other, so a chain
    is sequential; parallel width comes from running many chains (models, targets) at once."""
    L = len(seq)
    npz0 = {k: np.asarray(v) for k, v in np.load(npz_path).items()}
    counter = {"k": 0, "calls": 0}
    written = []

    def fold_fn(npz, n):
        counter["calls"] += 1
        return sampler.fold(ctx, [npz], seq, [n], seed=seed + 7919 * m + counter["calls"], params=params, rule=rule)

    def on_decoy(tag, xyz):
        counter["k"] += 1
        p = os.path.join(out_dir, "conf_%d_%d.pdb" % (m, counter["k"]))
        pdbio.write_pdb(p, seq, xyz, ["source %s model %d" % (tag, m)])
        written.append(p)

    dynamics.generate(fold_fn, npz0, L, n_init=init_num, n_max=n_max, angle=angle, on_decoy=on_decoy, seq=seq, ctx=ctx)
    return written


def run_many_from_npz(jobs, save_dir, init_num=10, n_max=300, angle=True, device=0, seed=0, rule="H1", ctx=None, streams=None):
    """The dynamics loop for many chains at once.  jobs: [(name, fasta, [NMR npz] or [NMR npz, X-ray npz]), ...]
    (the reference's name_lst loop over run_single, run_inference.py:339-354, with --mult_two_models).  Every
    (target, model) pair is an independent chain; `streams` chains (default: all, at most 16) run concurrently,
    one CUDA stream and one host thread each -- a chain folds ONE decoy per iteration and leaves a B200 almost
    idle on its own.  A decoy's result depends only on (seed, model, iteration), not on what runs beside it.
    ctx: None (contexts are created), one context (chains run one after the other on it) or a list of
    contexts.  Returns {name: [PDB paths]}, conf_{model}_{k}.pdb in the order the reference numbers them."""
    import queue
    from concurrent.futures import ThreadPoolExecutor
    chains = []
    for name, fasta, npz_paths in jobs:
        seq = read_fasta(fasta)
        out_dir = os.path.join(save_dir, name, "pred_pdb")
        os.makedirs(out_dir, exist_ok=True)
        for m, path in enumerate(npz_paths, start=1):
            chains.append((name, seq, path, m, out_dir))
    own = ctx is None
    if own:
        n_ctx = max(1, min(streams or len(chains), len(chains), 16))
        ctxs = [capi.Context(device) for _ in range(n_ctx)]
    else:
        ctxs = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]
    free = queue.Queue()
    for c in ctxs:
        free.put(c)

    def work(ch):
        name, seq, path, m, out_dir = ch
        params = tables.load_params()
        params["USE_ORIENT"] = bool(angle)
        c = free.get()
        try:
            return _run_chain(c, seq, path, m, out_dir, init_num, n_max, angle, seed, rule, params)
        finally:
            free.put(c)

    if len(ctxs) == 1:
        parts = [work(ch) for ch in chains]
    else:
        with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
            parts = list(ex.map(work, chains))
    out = {}
    for ch, files in zip(chains, parts):
        out.setdefault(ch[0], []).extend(files)
    if own:
        for c in ctxs:
            c.close()
    return out


def run_single_from_npz(name, fasta, npz_paths, save_dir, init_num=10, n_max=300, angle=True, device=0, seed=0,
                        rule="H1", ctx=None, streams=None):
    """npz_paths: [NMR npz] or [NMR npz, X-ray npz] (--mult_two_models).  The models' chains run concurrently
    (see run_many_from_npz).  Returns the list of PDB paths."""
    return run_many_from_npz([(name, fasta, npz_paths)], save_dir, init_num, n_max, angle, device, seed, rule, ctx, streams)[name]


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
    ap.add_argument("--streams", type=int, default=None, help="chains (models) in flight at once; default all")
    a = ap.parse_args(argv)
    files = run_single_from_npz(a.name, a.fasta, a.npz, a.save_dir, a.init_num, a.Nmax, a.angle, a.device, a.seed, streams=a.streams)
    print("wrote %d decoys under %s" % (len(files), os.path.join(a.save_dir, a.name, "pred_pdb")))


def fold_batch(ctx, targets, n_decoys, rank=0, world=1, seed=0, out_dir=None, params=None, rule="H1"):
    """Batch mode (the reference's name_lst loop, run_inference.py:339-354; BASELINE config 5): many
    targets of different length, n_decoys[t] decoys each, target-and-decoy sharded over `world` ranks.
    targets: [(name, seq, [npz, ...]), ...]; decoys of a target are dealt to its npz models in equal
    contiguous shares.  Every rank derives the same assignment (parallel.assign_blocks) and folds its
    blocks; a decoy's start (and therefore its result, bit for bit) depends only on (seed, target, global
    decoy index), not on the sharding.  ctx: one context, or a list of contexts (one CUDA stream each):
    that many targets are then in flight on the GPU at once, one host thread per context -- a batch of
    100 decoys of one target leaves most of a B200 idle.  Returns {(t, decoy): dict(xyz, tors, terms,
    model)} for this rank's decoys and writes out_dir/name/initial{decoy}.pdb when out_dir is given."""
    import queue
    from concurrent.futures import ThreadPoolExecutor
    from . import parallel
    params = params or tables.load_params()
    ctxs = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]
    plan = parallel.assign_blocks([len(t[1]) for t in targets], n_decoys, world)[rank]
    by_target = {}
    for t, d0, cnt in plan:
        by_target.setdefault(t, []).append((d0, cnt))
    free = queue.Queue()
    for c in ctxs:
        free.put(c)

    def fold_target(t):
        name, seq, npzs = targets[t]
        L, n_t, nm = len(seq), int(n_decoys[t]), len(npzs)
        c = free.get()
        out_t = {}
        try:
            tabs = [sampler.build_tables(c, z, seq, params, rule=rule) for z in npzs]
            starts = sampler.random_torsions(n_t, L, seed + 104729 * t)
            model = np.minimum(np.arange(n_t) * nm // max(n_t, 1), nm - 1)
            ids = np.concatenate([np.arange(d0, d0 + cnt) for d0, cnt in by_target[t]])
            for m in range(nm):
                sel = ids[model[ids] == m]
                if len(sel) == 0:
                    continue
                batch = capi.FoldBatch(c, [tabs[m]], [len(sel)], sampler.aa_index(seq), schedule_for(params))
                out = batch.run(starts[sel])
                batch.close()
                for k, d in enumerate(sel):
                    out_t[(t, int(d))] = dict(xyz=out["xyz"][k], tors=out["tors"][k], terms=out["terms"][k], model=m)
                    if out_dir:
                        p = os.path.join(out_dir, name, "initial%d.pdb" % d)
                        os.makedirs(os.path.dirname(p), exist_ok=True)
                        pdbio.write_pdb(p, seq, out["xyz"][k], ["target %s decoy %d model %d" % (name, d, m)])
            for tb in tabs:
                tb.close()
        finally:
            free.put(c)
        return out_t

    results = {}
    order = sorted(by_target, key=lambda t: -len(targets[t][1]))   # longest targets first: the tail is made of short ones
    if len(ctxs) == 1:
        for t in order:
            results.update(fold_target(t))
    else:
        with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
            for part in ex.map(fold_target, order):
                results.update(part)
    return results


def schedule_for(params):
    from . import schedule
    return schedule.reference_schedule()


if __name__ == "__main__":
    main()
