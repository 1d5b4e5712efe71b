"""Timings of the other BASELINE.json configurations (parity-test cases, not bench lines), for the record."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, metrics, sampler, schedule, synth, tables
ctx = capi.Context(0)
params = tables.load_params()

def fold(L, n, seed, dist_only=False, two=False, **kw):
    seq, npzs, nat = synth.target(L, seed=seed, two_model=two, **kw)
    zs = [{"dist": z["dist"]} for z in npzs] if dist_only else npzs
    tabs = [sampler.build_tables(ctx, z, seq, params) for z in zs]
    nd = [n] if len(tabs) == 1 else [n // 2, n - n // 2]
    b = capi.FoldBatch(ctx, tabs, nd, sampler.aa_index(seq), schedule.reference_schedule())
    t0 = sampler.random_torsions(n, L, seed)
    b.run(t0[:n])
    t = time.perf_counter(); out = b.run(t0); dt = time.perf_counter() - t
    tm = np.median([metrics.tm_score(c, nat[:, 1]) for c in out["xyz"][:16, :, 1].astype(np.float64)])
    b.close()
    for tb in tabs:
        tb.close()
    return dt, out["evals"].mean(), tm

dt, ev, tm = fold(150, 256, 150, dist_only=True)
print("config 2: L=150 distance-only, 256 decoys: %.2f s (%.0f decoys/s), %.0f evals/decoy, median TM %.2f" % (dt, 256 / dt, ev, tm))
dt, ev, tm = fold(800, 512, 800)   # (the CA-trace generator of synth.target(protein_like=False) takes minutes at L=800: not used here)
print("config 4: L=800 multi-domain, 512 decoys (no MC): %.2f s (%.0f decoys/s), %.0f evals/decoy, median TM %.2f" % (dt, 512 / dt, ev, tm))
