import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, synth, tables
L = int(os.environ.get("AB_L", "48"))
seq, npzs, nat = synth.target(L, seed=300 if L == 300 else 12, two_model=True)
ctx = capi.Context(0)
params = tables.load_params()
tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
aa = sampler.aa_index(seq)
nd = [160, 75] if L == 48 else [512, 512]
t0 = sampler.random_torsions(sum(nd), L, seed=4)
res = {}
for flag in ("1", "0"):
    os.environ["TRX_NO_K1SKIP"] = flag
    for nr in ((5, 8, 14) if L == 48 else (14,)):
        runs = schedule.reference_schedule()[:nr]
        for r in runs:
            r.skip_to = min(r.skip_to, nr)
        batch = capi.FoldBatch(ctx, tabs, nd, aa, runs)
        res[(flag, nr)] = batch.run(t0)
        batch.close()
for nr in ((5, 8, 14) if L == 48 else (14,)):
    a, b = res[("1", nr)], res[("0", nr)]
    for key in ("tors", "terms", "evals", "iters"):
        d = np.nonzero(np.any((a[key] != b[key]).reshape(len(t0), -1), axis=1))[0]
        print(nr, key, "differ:", len(d), d[:10])
    if nr == 5:
        d = np.nonzero(np.any((a["tors"] != b["tors"]).reshape(len(t0), -1), axis=1))[0]
        for n in d[:3]:
            print(" decoy", n, "evals", a["evals"][n], b["evals"][n], "terms", a["terms"][n], b["terms"][n])
