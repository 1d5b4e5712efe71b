"""Development: fold the reference's example target on the GPU and print decoy quality."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trx2dyn
from trx2dyn import capi, metrics, sampler, synth, tables
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = capi.Context(0)
ctx.set_timing(True)
if len(sys.argv) > 2:
    L = int(sys.argv[2]); seq, npzs, natx = synth.target(L, L, two_model=True)
    nat = {"apo": natx[:, 1], "holo": natx[:, 1]}
else:
    seq = open(f"{g}/example_seq.fasta").read().split("\n")[1]
    npzs = [np.load(f"{g}/example_NMR.npz"), np.load(f"{g}/example_Xray.npz")]
    nat = np.load(f"{g}/example_natives_ca.npz")
t = time.time(); out = sampler.fold(ctx, npzs, seq, [n // 2, n // 2], seed=1); dt = time.time() - t
print("decoys", n, "L", len(seq), "wall %.2fs -> %.1f decoys/s" % (dt, n / dt), "rounds", out["rounds"])
print("evals: mean %.0f max %d ; iters mean %.0f" % (out["evals"].mean(), out["evals"].max(), out["iters"].mean()))
print("terms median", np.round(np.median(out["terms"], 0), 1))
ca = out["xyz"][:, :, 1].astype(np.float64)
k = min(n, 32)
tma = np.array([metrics.tm_score(c, nat["apo"]) for c in ca[:k]]); tmh = np.array([metrics.tm_score(c, nat["holo"]) for c in ca[:k]])
print("TM apo", np.round(np.sort(tma), 3)); print("TM holo", np.round(np.sort(tmh), 3))
for name in ("nerf", "restraints", "reduce", "centroid", "torsion_grad", "lbfgs", "activity"):
    ms, cnt = ctx.timing(name); print("%-13s %8.2f ms %6d launches %.1f us each" % (name, ms, cnt, 1e3 * ms / max(cnt, 1)))
