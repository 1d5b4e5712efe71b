#!/usr/bin/env python
"""bench.py -- headline benchmark of the folding hot path (contract: see DESIGN.md 'Measurement').

Workload (BASELINE.json configs[2]): synthetic L=300 target, full dist+omega+theta+phi
restraints, two-model mixing (half the decoys scored against each model's tables),
4096 decoys per GPU, decoys sharded across GPUs with no data-path collective (weak scaling).

One "step" = one pass of the hot path over the whole decoy batch.  With --mode restraint
(stage available in every build) the pass is one restraint energy+gradient evaluation of
every decoy (SURVEY 8d metric M2); with --mode fold it is a complete centroid fold of
every decoy (metric M1, decoys/s).  The default is the most complete mode the library has.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode ...]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

L_TARGET = 300
SEED = 300
WEIGHTS = (5.0, 4.0, 4.0)  # folding/data/scorefxn.wts: atom_pair 5, dihedral 4, angle 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fold", choices=["restraint", "fold"])
    ap.add_argument("--decoys", type=int, default=4096, help="decoys per GPU")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lbfgs-m", type=int, default=20)
    ap.add_argument("--streams", type=int, default=1, help="independent fold batches per GPU (own stream each)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        late = False
        t0 = time.time()
        while not self.rows and time.time() - t0 < 1.5:   # timed region shorter than the 100 ms sampling period
            late = True
            time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 8 and r[4 + k].lower().startswith("active") for r in self.rows)]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm)}
        if late:
            out["note"] = "timed region shorter than the sampling period: first sample taken right after it"
        return out


def build_workload(n_decoys, precision):
    """Two table sets (two-model mixing) + decoy coordinates near the synthetic native."""
    import trx2dyn  # noqa: F401
    from trx2dyn import synth, tables
    seq, npzs, nat = synth.target(L_TARGET, SEED, dense=False, two_model=True)
    params = tables.load_params()
    acts = []
    for npz in npzs:
        rst = tables.gen_rst(npz, params)
        acts.append(tables.active_restraints(rst, tables.select(rst, 1, L_TARGET, params)))
    rng = np.random.default_rng(SEED + 7)
    sig = rng.uniform(0.3, 3.0, size=(n_decoys, 1, 1, 1))
    xyz = nat[None, :, [0, 1, 3]] + rng.normal(size=(n_decoys, L_TARGET, 3, 3)) * sig
    return seq, npzs, acts, xyz.astype(np.float32 if precision == 32 else np.float64)


def oracle_sets(npzs):
    from oracle.tables_oracle import gen_rst_oracle, select_oracle
    from oracle import restraints_oracle as ro
    sets = []
    for npz in npzs:
        rst = gen_rst_oracle(npz)
        sets.append(ro.RestraintSetOracle(rst, select_oracle(rst, 1, L_TARGET, 0.05), "H1"))
    return sets


def cpu_restraint_rate(npzs, xyz, n_sample, threads):
    """Oracle (CPU port of the same arithmetic) timed on a bounded sample of the workload."""
    sets = oracle_sets(npzs)
    half = n_sample // 2
    xs = [np.ascontiguousarray(xyz[:half], dtype=np.float64), np.ascontiguousarray(xyz[half:n_sample], dtype=np.float64)]
    sets[0].energy_grad_batch(xs[0][:threads], WEIGHTS, threads)  # warm
    t0 = time.perf_counter()
    for s, x in zip(sets, xs):
        s.energy_grad_batch(x, WEIGHTS, threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def cpu_fold_rate(npzs, seq, n_sample, threads, seed):
    """Oracle fold (same schedule, fp64, one decoy per host thread) on a bounded sample."""
    from oracle import fold_oracle as fo
    sets = oracle_sets(npzs)
    half = max(1, n_sample // 2)
    t0 = time.perf_counter()
    evals = 0
    for k, rs in enumerate(sets):
        F = fo.FoldOracle(rs, seq)
        n = half if k == 0 else n_sample - half
        if n <= 0:
            continue
        out = F.fold(fo.random_torsions(n, len(seq), seed + k), fo.reference_schedule(), m=20, nthreads=threads)
        evals += int(out["evals"].sum())
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, evals


def run_reference_fold(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps or 2
    n_sample = threads
    import trx2dyn  # noqa: F401
    from trx2dyn import synth
    seq, npzs, _ = synth.target(L_TARGET, SEED, dense=False, two_model=True)
    t = []
    for k in range(min(args.warmup, 1) + steps):
        rate, dt, _ = cpu_fold_rate(npzs, seq, n_sample, threads, SEED + 10 * k)
        if k >= min(args.warmup, 1):
            t.append(dt)
    val = n_sample / float(np.mean(t))
    sample = "%d decoys (one per host thread) of the L=300 two-model workload per step, oracle/fold_oracle.c, same schedule, fp64 (PyRosetta absent)" % n_sample
    line = {"impl": "reference", "metric": "decoys_per_sec_L300", "value": val, "unit": "decoys/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic L=300 dist+omega+theta+phi, two-model mixing (configs[2])", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "decoys/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "decoys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  PyRosetta is absent
    (un-vendored binary dependency), so this times the oracle port on all host threads."""
    if args.mode != "restraint":
        return run_reference_fold(args)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps or 3
    n_sample = max(2 * threads, 64)
    _, npzs, _, xyz = build_workload(n_sample, 64)
    for _ in range(min(args.warmup, 1)):
        cpu_restraint_rate(npzs, xyz, n_sample, threads)
    t = []
    for _ in range(steps):
        t.append(cpu_restraint_rate(npzs, xyz, n_sample, threads)[1])
    val = n_sample / float(np.mean(t))
    line = {"impl": "reference", "metric": "restraint_energy_grad_decoy_evals_per_sec_L300", "value": val,
            "unit": "decoy-evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic L=300 dist+omega+theta+phi, two-model mixing (configs[2])",
                       "sample": "%d decoys per step" % n_sample},
            "cpu_baseline": {"value": val, "unit": "decoy-evals/s", "cores": threads, "kind": "port",
                             "sample": "%d decoys of the L=300 workload per step, oracle/restraints_oracle.c on %d pthreads (PyRosetta absent)" % (n_sample, threads)},
            "e2e": {"value": val, "unit": "decoy-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_b200_fold(args):
    """Metric M1: fully minimised centroid decoys per second (whole schedule, on device)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import trx2dyn  # noqa: F401
    from trx2dyn import capi, sampler, schedule, synth, tables

    N = args.decoys
    steps = args.steps or 3
    warmup = max(args.warmup, 3)
    seq, npzs, nat = synth.target(L_TARGET, SEED, dense=False, two_model=True)
    params = tables.load_params()
    # The batch is split over `streams` independent fold batches (own context + CUDA stream each,
    # driven by one host thread each): while one batch is in its latency-bound tail (few decoys
    # left) the others keep the SMs busy.  Same work, same results per decoy.
    S = max(1, args.streams)
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    per = [(N // S + (1 if k < N % S else 0)) for k in range(S)]
    offs = np.concatenate([[0], np.cumsum(per)]).astype(int)
    lanes = []
    for k in range(S):
        stream = torch.cuda.Stream()
        ctx_k = capi.Context(local, stream.cuda_stream)
        tabs_k = [sampler.build_tables(ctx_k, npz, seq, params) for npz in npzs]
        half = (per[k] // 2 + 31) // 32 * 32
        nd_k = [half, per[k] - half]
        batch_k = capi.FoldBatch(ctx_k, tabs_k, nd_k, sampler.aa_index(seq), schedule.reference_schedule(), lbfgs_m=args.lbfgs_m)
        lanes.append(dict(stream=stream, ctx=ctx_k, tabs=tabs_k, nd=nd_k, batch=batch_k, rounds=C.c_int()))
    R = [sum(t.info()["counts"]) for t in lanes[0]["tabs"]]
    pool_exec = ThreadPoolExecutor(max_workers=S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host buffers: torsions in, coordinates + terms out (the e2e path IS the product call)
    tors_h = torch.empty((N, L_TARGET, 3), dtype=torch.float32).pin_memory()
    xyz_h = torch.empty((N, L_TARGET, 5, 3), dtype=torch.float32).pin_memory()
    terms_h = torch.empty((N, 7), dtype=torch.float64).pin_memory()
    stats_h = torch.empty((N, 2), dtype=torch.int64).pin_memory()

    def fold_lane(k):
        ln, o, n = lanes[k], int(offs[k]), per[k]
        capi.check(capi.lib().trx_fold_run(ln["batch"]._h, C.c_void_p(tors_h[o:o + n].data_ptr()), C.c_void_p(xyz_h[o:o + n].data_ptr()),
                                           C.c_void_p(terms_h[o:o + n].data_ptr()), C.c_void_p(stats_h[o:o + n].data_ptr()),
                                           C.c_int(20000), C.c_int(16), C.byref(ln["rounds"])))

    def one_fold(seed):
        tors_h.copy_(torch.from_numpy(sampler.random_torsions(N, L_TARGET, seed)))
        t0 = time.perf_counter()
        list(pool_exec.map(fold_lane, range(S)))
        return time.perf_counter() - t0

    ctx = lanes[0]["ctx"]
    rounds = lanes[0]["rounds"]
    nd = None

    for k in range(warmup):
        one_fold(1000 * rank + k)
    barrier()
    sampler_clk = ClockSampler(local)
    sampler_clk.start()
    for ln in lanes:
        ln["ctx"].set_timing(True)
        ln["ctx"].reset_timing()
    launches0 = sum(ln["ctx"].launch_count for ln in lanes)
    barrier()
    from trx2dyn import parallel
    t_wall, evals_total, rest_evals = [], 0, 0.0
    pool = None
    for k in range(steps):
        t0 = time.perf_counter()
        dt = one_fold(1000 * rank + 100 + k)
        if world > 1:
            # the path's only exchange: all-gather of per-decoy energies (NCCL) for pool selection
            score = (terms_h.numpy() * np.array([5.0, 4.0, 4.0, 1.0, 1.0, 0.5, 0.1])).sum(1)
            full = parallel.gather_scalars(score, N * world, rank, world, device="cuda")
            pool = parallel.select_pool(full[:, 0], 10)
            dt = time.perf_counter() - t0
        t_wall.append(dt)
        ev = stats_h[:, 0].numpy().astype(np.float64)
        evals_total += float(ev.sum())
        for q in range(S):
            o, n0 = int(offs[q]), lanes[q]["nd"][0]
            rest_evals += float(ev[o:o + n0].sum()) * R[0] + float(ev[o + n0:int(offs[q + 1])].sum()) * R[1]
    barrier()
    clocks = sampler_clk.stop()
    launches = sum(ln["ctx"].launch_count for ln in lanes) - launches0
    # device time of a step: the slowest lane's fold (CUDA events on its stream, inputs resident);
    # lanes run concurrently, so steps cost max over lanes, not the sum
    lane_dev = [ln["ctx"].timing("fold_device")[0] / 1e3 for ln in lanes]
    t_dev = max(lane_dev)
    k1_ms = sum(ln["ctx"].timing("restraints")[0] for ln in lanes)
    k1_n = sum(ln["ctx"].timing("restraints")[1] for ln in lanes)
    busy = {name: sum(ln["ctx"].timing(name)[0] for ln in lanes) for name in ("restraints", "reduce", "nerf", "centroid", "torsion_grad", "lbfgs", "cart_gather", "cart_grad", "segment", "compact", "activity")}
    tot_busy = sum(busy.values())
    shares = {name: v / tot_busy for name, v in busy.items()}
    counts = {name: sum(ln["ctx"].timing(name)[1] for ln in lanes) for name in busy}
    for ln in lanes:
        ln["ctx"].set_timing(False)
    t_e2e = float(sum(t_wall))
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ca = xyz_h[: min(N, 64), :, 1].numpy().astype(np.float64)
    from trx2dyn import metrics
    tm = [metrics.tm_score(c, nat[:, 1]) for c in ca[:32]]
    total = N * world
    peak, peak_src = peaks()
    # algorithmic bytes the restraint kernel processed: 16 B (4 fp32 knot scalars) per
    # restraint evaluated for a decoy + coordinates in / gradient out per decoy evaluation
    alg_bytes = 16.0 * rest_evals + 2 * 3 * L_TARGET * 12.0 * evals_total
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    line = {"metric": "decoys_per_sec_L300", "value": total * steps / t_dev, "unit": "decoys/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic L=300 dist+omega+theta+phi, two-model mixing, %d decoys per GPU, full mode-2 centroid schedule (configs[2])" % N,
                       "restraints_per_decoy": R, "l2": "working set per step (%.1f GB of decoy state) exceeds L2" % (batch_bytes(N, L_TARGET, args.lbfgs_m) / 1e9),
                       "streams": S,
                       "mode": "fold", "lbfgs_m": args.lbfgs_m, "cartesian_stage": "min_mover_cart built: coordinates as degrees of freedom, cart_bonded-like springs (stated approximation; hbond_* terms dropped)",
                       "collective": "all-gather of per-decoy energies for pool selection (N>1 only)"},
            "restraint_decoy_evals_per_sec": evals_total * world / t_dev,
            "restraint_evals_per_sec": rest_evals * world / t_dev,
            "mean_evals_per_decoy": evals_total / (N * steps), "rounds_last_step": rounds.value,
            "decoy_quality": {"tm_vs_synthetic_native_median": float(np.median(tm)), "tm_gt_0.5_frac": float(np.mean(np.array(tm) > 0.5))},
            "clocks": clocks,
            "e2e": {"value": total * steps / t_e2e, "unit": "decoys/s",
                    "h2d_bytes_per_step": int(tors_h.numel() * 4),
                    "d2h_bytes_per_step": int(xyz_h.numel() * 4 + tors_h.numel() * 4 + terms_h.numel() * 8 + stats_h.numel() * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "restraints_kernel<float>", "kernel_ms": k1_ms / max(k1_n, 1),
                         "kernel_share_of_step": shares["restraints"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / max(k1_n, 1), "kernel_shares": shares,
                         "kernel_ms_total": {k: round(v, 2) for k, v in busy.items()}, "kernel_launches": counts}}
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only (the other ranks of an N>1 run would wait on it)
        threads = os.cpu_count() or 1
        rate, dt, ev = cpu_fold_rate(npzs, seq, threads, threads, SEED)
        line["cpu_baseline"] = {"value": rate, "unit": "decoys/s", "cores": threads, "kind": "port",
                                "sample": "%d decoys (one per host thread) of the same workload, oracle/fold_oracle.c same schedule fp64, %.1f s (PyRosetta absent)" % (threads, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def batch_bytes(N, L, m):
    return N * L * 3 * 4.0 * (5 + 2 * m) + N * L * 15 * 4.0 * 3


def run_b200(args):
    if args.mode != "restraint":
        return run_b200_fold(args)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import trx2dyn  # noqa: F401
    from trx2dyn import capi

    N, prec = args.decoys, args.precision
    steps = args.steps or 10
    seq, npzs, acts, xyz = build_workload(N, prec)
    # rank r scores its own shard: different decoys per rank, same tables
    if world > 1:
        rng = np.random.default_rng(1000 + rank)
        xyz = (xyz + rng.normal(size=xyz.shape).astype(xyz.dtype) * 0.05).astype(xyz.dtype)
    stream = torch.cuda.Stream()
    ctx = capi.Context(local, stream.cuda_stream)
    tabs = [capi.Tables(ctx, L_TARGET, a) for a in acts]
    R = [sum(t.info()["counts"]) for t in tabs]
    half = N // 2
    parts = [(0, half), (half, N - half)]
    tdt = torch.float32 if prec == 32 else torch.float64
    es = prec // 8
    Lpad = capi.padded_length(L_TARGET)

    dev = []
    with torch.cuda.stream(stream):
        for (o, n) in parts:
            nat = torch.tensor(xyz[o:o + n], device="cuda")
            G = (n + 31) // 32
            grp = torch.empty(G * Lpad * 9 * 32, dtype=tdt, device="cuda")
            capi.to_grouped(ctx, n, L_TARGET, 3, prec, nat.data_ptr(), grp.data_ptr())
            dev.append(dict(n=n, grp=grp, grad=torch.empty_like(grp), E=torch.empty(3 * G * 32, dtype=torch.float64, device="cuda")))
        flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    ctx.sync()

    def step_device():
        for tb, d in zip(tabs, dev):
            tb.energy_grad_device(d["n"], d["grp"].data_ptr(), d["E"].data_ptr(), d["grad"].data_ptr(), WEIGHTS, prec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value): CUDA events on the launching stream, L2 flushed between steps
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.set_timing(True)
    ctx.reset_timing()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            step_device()
            b.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    k1_ms, k1_n = ctx.timing("restraints")
    ctx.set_timing(False)
    if world > 1:
        t = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    host = [torch.tensor(xyz[o:o + n]).pin_memory() for (o, n) in parts]
    hostE = [torch.empty((n, 3), dtype=torch.float64).pin_memory() for (_, n) in parts]
    hostG = [torch.empty_like(h).pin_memory() for h in host]
    import ctypes as C
    w = np.ascontiguousarray(WEIGHTS, dtype=np.float64)

    def step_e2e():
        for tb, h, e, g in zip(tabs, host, hostE, hostG):
            capi.check(capi.lib().trx_energy_grad(ctx._h, tb._h, C.c_int(h.shape[0]), C.c_int(prec), C.c_void_p(h.data_ptr()),
                                                  w.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(e.data_ptr()),
                                                  C.c_void_p(g.data_ptr())))
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(3, steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total = N * world
    value = total * steps / t_dev
    peak, peak_src = peaks()
    # algorithmic bytes of one restraint-kernel launch: 4 knot scalars per restraint-eval
    # + coordinates read + gradient written, per decoy; launches alternate between the two table sets
    alg = [(4 * es * r + 2 * 3 * L_TARGET * 3 * es) * n for r, (_, n) in zip(R, parts)]
    alg_per_launch = float(np.mean(alg))
    k1_avg_ms = k1_ms / max(k1_n, 1)
    achieved = alg_per_launch / (k1_avg_ms * 1e-3) / 1e9
    line = {"metric": "restraint_energy_grad_decoy_evals_per_sec_L300", "value": value, "unit": "decoy-evals/s",
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if prec == 32 else "f64", "data": "synthetic",
            "config": {"workload": "synthetic L=300 dist+omega+theta+phi, two-model mixing, %d decoys per GPU (configs[2])" % N,
                       "restraints_per_decoy": R, "l2": "flushed between steps (160 MB write)", "mode": "restraint"},
            "restraint_evals_per_sec": value * float(np.mean(R)),
            "clocks": clocks,
            "e2e": {"value": total * e2e_steps / t_e2e, "unit": "decoy-evals/s",
                    "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host)),
                    "d2h_bytes_per_step": int(sum(g.numel() * g.element_size() for g in hostG) + sum(e.numel() * 8 for e in hostE))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "restraints_kernel<float>" if prec == 32 else "restraints_kernel<double>",
                         "kernel_ms": k1_avg_ms, "kernel_share_of_step": k1_ms / (1e3 * t_dev), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_per_launch}}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        n_sample = max(2 * threads, 64)
        rate, dt = cpu_restraint_rate(npzs, xyz, n_sample, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "decoy-evals/s", "cores": threads, "kind": "port",
                                "sample": "%d decoys of the same workload, oracle/restraints_oracle.c on %d pthreads, %.2f s (PyRosetta absent)" % (n_sample, threads, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
