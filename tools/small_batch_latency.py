"""Wall time of small folds (1, 32, 256 decoys) on the reference's example target: the regime of the
drop-in CLI (one decoy per call) and of the outer dynamics loop (one decoy per iteration)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, sampler, schedule, tables
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
seq = open(f"{g}/example_seq.fasta").read().split("\n")[1]
npz = np.load(f"{g}/example_NMR.npz")
ctx = capi.Context(0)
tb = sampler.build_tables(ctx, npz, seq, tables.load_params())
for n in (1, 32, 256):
    batch = capi.FoldBatch(ctx, [tb], [n], sampler.aa_index(seq), schedule.reference_schedule())
    t0 = sampler.random_torsions(n, len(seq), 1)
    batch.run(t0)
    ts = []
    for k in range(3):
        t = time.perf_counter(); out = batch.run(t0); ts.append(time.perf_counter() - t)
    print("decoys %4d  fold %.3f s  rounds %d  launches/round ~%.1f  us/round %.1f" % (n, min(ts), out["rounds"], 12, 1e6 * min(ts) / out["rounds"]))
    batch.close()
