"""Where one iteration of the dynamics loop spends its host time (example target, one chain)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, dynamics, sampler, schedule, tables

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
seq = open(os.path.join(G, "example_seq.fasta")).read().split("\n")[1].strip()
L = len(seq)
npz = {k: np.asarray(v) for k, v in np.load(os.path.join(G, "example_NMR.npz")).items()}
ctx = capi.Context(0)
params = tables.load_params()
state = capi.DynState(ctx, npz)
T = {}
def tick(name, t0):
    ctx.sync()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
cur = npz
for it in range(12):
    t0 = time.perf_counter(); rst = tables.gen_rst(cur, params, use_orient=True); tick("gen_rst (numpy)", t0)
    t0 = time.perf_counter(); masks = tables.select(rst, 1, L, params, seq, False); act = tables.active_restraints(rst, masks, "H1"); tick("select + knots (numpy)", t0)
    t0 = time.perf_counter(); tb = capi.Tables(ctx, L, act); tick("Tables create (spline fit, tiles, schedule)", t0)
    t0 = time.perf_counter(); batch = capi.FoldBatch(ctx, [tb], [1], sampler.aa_index(seq), schedule.reference_schedule()); tick("FoldBatch create", t0)
    t0 = time.perf_counter(); out = batch.run(sampler.random_torsions(1, L, it)); tick("fold (1 decoy)", t0)
    rounds = out["rounds"]
    t0 = time.perf_counter(); batch.close(); tb.close(); tick("close", t0)
    xyz = out["xyz"][0].astype(np.float64)
    t0 = time.perf_counter(); state.step(xyz[:, 0], xyz[:, 1], xyz[:, 3], cb=xyz[:, 2], seq=seq); cur = state.get(); tick("distogram update (device) + download", t0)
    if it == 1:
        T = {}          # the first two iterations warm everything up
print("per iteration over 10 iterations (ms); last fold: %d rounds, %d evaluations" % (rounds, out["evals"][0]))
for k, v in T.items():
    print("  %-46s %7.1f" % (k, 1e3 * v / 10))
print("  %-46s %7.1f" % ("total", 1e3 * sum(T.values()) / 10))
