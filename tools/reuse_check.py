"""A fold batch reused for a second fold must give what a fresh batch gives."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, synth, tables
L, N = 300, int(os.environ.get("DET_N", "2048"))
seq, npzs, nat = synth.target(L, 300, dense=False, two_model=True)
ctx = capi.Context(0)
params = tables.load_params()
tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
aa = sampler.aa_index(seq)
half = (N // 2 + 31) // 32 * 32
tA, tB = sampler.random_torsions(N, L, 1), sampler.random_torsions(N, L, 2)
b = capi.FoldBatch(ctx, tabs, [half, N - half], aa, schedule.reference_schedule())
b.run(tA)
reused = b.run(tB)
b.close()
b = capi.FoldBatch(ctx, tabs, [half, N - half], aa, schedule.reference_schedule())
fresh = b.run(tB)
b.close()
for key in ("tors", "xyz", "terms", "evals", "iters"):
    bad = np.nonzero(np.any((reused[key] != fresh[key]).reshape(N, -1), axis=1))[0]
    print(key, "differing decoys:", len(bad), bad[:10])
print("evals", reused["evals"].mean(), fresh["evals"].mean(), "rounds", reused["rounds"], fresh["rounds"])
