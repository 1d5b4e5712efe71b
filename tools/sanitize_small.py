"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): tables, K1 fp32+fp64, the fold with its
Cartesian segment, MC, Cartesian evaluation entry, decoy-set metrics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, synth, tables
L = 40
seq, npzs, nat = synth.target(L, seed=3, two_model=True)
ctx = capi.Context(0)
params = tables.load_params()
tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
xyz = synth.random_backbones(37, L, 1)
for prec in (capi.F64, capi.F32):
    E, g = tabs[0].energy_grad(xyz, (5, 4, 4), prec)
    assert np.all(np.isfinite(E)) and np.all(np.isfinite(g))
runs = schedule.mc_schedule(mc_max_iter=20)
for r in runs:
    r.max_iter = min(r.max_iter, 8)
batch = capi.FoldBatch(ctx, tabs, [32, 9], sampler.aa_index(seq), runs, lbfgs_m=8)
out = batch.run_mc(sampler.random_torsions(41, L, 0), cycles=1, kT=1.0, seed=1, max_rounds=120)
assert np.all(np.isfinite(out["terms"]))
assert [r.cartesian for r in runs].count(1) == 1
tot, terms, grad, tors = batch.eval_cart(out["xyz"], np.array(list(runs[8].w)))
assert np.all(np.isfinite(grad)) and np.all(np.isfinite(tors))
from trx2dyn import metrics
tm, rm = metrics.tmscore_matrix(ctx, out["xyz"][:9, :, 1])
gl = metrics.glocon_matrix(ctx, out["xyz"][:9, :, 2])
assert np.all(np.isfinite(tm)) and np.all(np.isfinite(rm)) and np.all(np.isfinite(gl))
print("sanitize_small ok", out["rounds"], out["evals"].mean())
