"""Bench-sized determinism check: the same 4096-decoy two-model fold twice, and once more with the
restraint-kernel skip / decoy packing switched off; all outputs must be bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, synth, tables
L, N = 300, int(os.environ.get("DET_N", "4096"))
seq, npzs, nat = synth.target(L, 300, dense=False, two_model=True)
ctx = capi.Context(0)
params = tables.load_params()
tabs = [sampler.build_tables(ctx, z, seq, params) for z in npzs]
aa = sampler.aa_index(seq)
half = (N // 2 + 31) // 32 * 32
t0 = sampler.random_torsions(N, L, 1100)
def run(**env):
    for k, v in env.items():
        os.environ[k] = v
    b = capi.FoldBatch(ctx, tabs, [half, N - half], aa, schedule.reference_schedule())
    out = b.run(t0)
    b.close()
    return out
ref = run(TRX_NO_K1SKIP="0", TRX_NO_MIGRATE="0")
print("evals/decoy", ref["evals"].mean(), "rounds", ref["rounds"])
for name, env in (("again", {}), ("no k1 skip", {"TRX_NO_K1SKIP": "1"}), ("no migration", {"TRX_NO_K1SKIP": "0", "TRX_NO_MIGRATE": "1"})):
    o = run(**env)
    bad = np.nonzero(np.any((o["tors"] != ref["tors"]).reshape(N, -1), axis=1))[0]
    print("%-14s differing decoys: %d %s  evals/decoy %.4f rounds %d" % (name, len(bad), bad[:8], o["evals"].mean(), o["rounds"]))
