"""How the energy evaluations of a fold split over the schedule: prefixes of the schedule on the bench target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, synth, tables
L = 300
seq, npzs, nat = synth.target(L, 300, dense=False, two_model=True)
ctx = capi.Context(0)
tb = sampler.build_tables(ctx, npzs[0], seq, tables.load_params())
t0 = sampler.random_torsions(512, L, 1100)
full = schedule.reference_schedule()
prev = 0.0
for name, n in (("remove_clash(vdw) x5", 5), ("+ min_mover x3", 8), ("+ min_mover_cart", 9), ("+ remove_clash(min_mover1) x5", 14)):
    runs = full[:n]
    for r in runs:
        if r.skip_to > n:
            r.skip_to = n
    batch = capi.FoldBatch(ctx, [tb], [512], sampler.aa_index(seq), runs)
    out = batch.run(t0)
    batch.close()
    ev = out["evals"].mean()
    print("%-32s evals/decoy %7.1f  (+%.1f)  rounds %d" % (name, ev, ev - prev, out["rounds"]))
    prev = ev
    full = schedule.reference_schedule()
