"""Development micro-benchmark of the restraint kernel alone (device-resident inputs)."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trx2dyn
from trx2dyn import capi, synth, tables

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=300)
ap.add_argument("--N", type=int, default=4096)
ap.add_argument("--dense", action="store_true")
ap.add_argument("--prec", type=int, default=32)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--native-like", action="store_true")
a = ap.parse_args()
seq, npzs, nat = synth.target(a.L, a.L, dense=a.dense)
params = tables.load_params()
rst = tables.gen_rst(npzs[0], params)
act = tables.active_restraints(rst, tables.select(rst, 1, a.L, params))
stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
tb = capi.Tables(ctx, a.L, act)
R = sum(tb.info()["counts"])
dt = torch.float32 if a.prec == 32 else torch.float64
if a.native_like:
    rng = np.random.default_rng(0)
    xyz = nat[None, :, [0, 1, 3]] + rng.normal(size=(a.N, a.L, 3, 3)) * 1.0
else:
    xyz = synth.random_backbones(a.N, a.L, 1)
Lpad = capi.padded_length(a.L); G = (a.N + 31) // 32
with torch.cuda.stream(stream):
    nat_d = torch.tensor(xyz, dtype=dt, device="cuda")
    grp = torch.empty(G * Lpad * 9 * 32, dtype=dt, device="cuda")
    grad = torch.empty_like(grp)
    E = torch.empty(3 * G * 32, dtype=torch.float64, device="cuda")
    capi.to_grouped(ctx, a.N, a.L, 3, a.prec, nat_d.data_ptr(), grp.data_ptr())
    for _ in range(3):
        tb.energy_grad_device(a.N, grp.data_ptr(), E.data_ptr(), grad.data_ptr(), (5, 4, 4), a.prec)
    ctx.sync(); ctx.set_timing(True)
    for _ in range(a.iters):
        tb.energy_grad_device(a.N, grp.data_ptr(), E.data_ptr(), grad.data_ptr(), (5, 4, 4), a.prec)
    ctx.sync()
ms, n = ctx.timing("restraints"); ms2, n2 = ctx.timing("reduce")
per = ms / n
es = a.prec // 8
alg_bytes = (2 * es * 2 * R + 2 * 3 * a.L * 3 * es) * a.N   # 4 knot scalars per restraint + coords in + grad out
print(json.dumps(dict(L=a.L, N=a.N, dense=a.dense, prec=a.prec, restraints=R, tiles=tb.info()["tiles"],
                      k1_ms=per, reduce_ms=ms2 / n2, decoy_evals_per_s=a.N / (per + ms2 / n2) * 1e3,
                      restraint_evals_per_s=a.N * R / per * 1e3, alg_GBs=alg_bytes / per / 1e6)))
