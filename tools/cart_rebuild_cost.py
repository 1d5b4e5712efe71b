"""What the ideal-geometry rebuild after min_mover_cart costs: the same decoys folded (a) through the whole
schedule -- the final remove_clash(min_mover1) runs rebuild the chain from the read-back torsions with ideal bonds -- and
(b) with the schedule cut after the Cartesian run (the decoy keeps its Cartesian coordinates).  Reports TM-score vs the
synthetic native for both, the CA RMSD between the two versions of each decoy, and the restraint score of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, metrics, sampler, schedule, synth, tables

L, N = 300, 256
seq, npzs, nat = synth.target(L, 300, two_model=False)
ctx = capi.Context(0)
tb = sampler.build_tables(ctx, npzs[0], seq, tables.load_params())
aa = sampler.aa_index(seq)
t0 = sampler.random_torsions(N, L, 7)
full = schedule.reference_schedule()
cut = full[:9]          # ... up to and including min_mover_cart
res = {}
for tag, runs in (("full schedule", full), ("cut after min_mover_cart", cut)):
    b = capi.FoldBatch(ctx, [tb], [N], aa, runs)
    res[tag] = b.run(t0)
    b.close()
w = np.array(list(full[5].w))
a, c = res["full schedule"], res["cut after min_mover_cart"]
for tag, o in res.items():
    tm = np.array([metrics.tm_score(x, nat[:, 1]) for x in o["xyz"][:64, :, 1].astype(np.float64)])
    bond = np.linalg.norm(o["xyz"][:, :, 1] - o["xyz"][:, :, 0], axis=-1)
    print("%-26s TM median %.3f (q10 %.3f, q90 %.3f)  restraint score median %.0f  |N-CA - 1.458| max %.4f  evals/decoy %.0f" %
          (tag, np.median(tm), np.quantile(tm, 0.1), np.quantile(tm, 0.9), np.median(o["terms"][:, :3] @ w[:3]), np.abs(bond - 1.458).max(), o["evals"].mean()))
rm = []
for k in range(64):
    P, Q = a["xyz"][k, :, 1].astype(np.float64), c["xyz"][k, :, 1].astype(np.float64)
    P, Q = P - P.mean(0), Q - Q.mean(0)
    U, S, Vt = np.linalg.svd(P.T @ Q)
    d = np.sign(np.linalg.det(U @ Vt))
    rm.append(np.sqrt(max(0.0, (P ** 2).sum() + (Q ** 2).sum() - 2 * (S[0] + S[1] + d * S[2])) / L))
print("CA RMSD between the two versions of a decoy: median %.2f A, max %.2f A" % (np.median(rm), np.max(rm)))
