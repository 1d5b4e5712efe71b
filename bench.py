#!/usr/bin/env python
"""bench.py -- benchmark of the folding hot path (contract: DESIGN.md 'Measurement').

Default workload = BASELINE.json configs[2]: synthetic L=300 target, dist+omega+theta+phi restraints,
two-model mixing (half the decoys scored against each model's tables), full mode-2 centroid schedule,
decoys sharded across GPUs with no data-path collective.  A "step" is one call of the public fold entry
point (trx_fold_run_queue) over one batch of random starts: `--decoys` decoys per GPU folded through
`--resident` positions (continuous batching: a position is refilled as soon as its decoy leaves the
schedule segment in progress).  Defaults for configs[2]: 24576 decoys per GPU and step through 12288 positions (measured on one B200 with
4096 / 8192 / 12288 / 16384 positions: 1563 / 1625 / 1658 / 1630 decoys/s; 12 GB of decoy state; the step is
sized so that the driver's --steps 20 --warmup 5 stays well inside its per-run limit).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--config 1|2|3|4] [--scaling weak|strong] [--decoys D] [--resident R] [--mode fold|restraint]

--config 1  configs[1]: L=150 distance-only restraints (--no-angle), 256 decoys
--config 2  configs[2] (default)
--config 3  configs[3]: L=800, 2048 decoys, Monte-Carlo perturbation cycles, decoy-sharded
--config 4  configs[4]: batch mode, 64 targets L=100..500 (name_lst), 100 decoys each, target-and-decoy sharded
--scaling strong: --decoys is the TOTAL over all GPUs (configs[2]/[3] read "N decoys sharded 1/2/4/8")
--mode restraint: the restraint kernel alone (SURVEY 8d metric M2)
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

WEIGHTS = (5.0, 4.0, 4.0)  # folding/data/scorefxn.wts: atom_pair 5, dihedral 4, angle 4

CONFIGS = {
    1: dict(tag="configs[1]", L=150, seed=150, two_model=False, dist_only=True, decoys=256, resident=256, mc=None,
            metric="decoys_per_sec_L150_dist_only", desc="synthetic L=150 distance-only restraints (--no-angle)"),
    2: dict(tag="configs[2]", L=300, seed=300, two_model=True, dist_only=False, decoys=24576, resident=12288, mc=None,
            metric="decoys_per_sec_L300", desc="synthetic L=300 dist+omega+theta+phi, two-model mixing"),
    3: dict(tag="configs[3]", L=800, seed=800, two_model=False, dist_only=False, decoys=2048, resident=2048,
            mc=dict(cycles=4, kT=2.0, block=(3, 9), sigma_deg=20.0, mc_max_iter=200),
            metric="decoys_per_sec_L800_mc", desc="synthetic L=800 multi-domain target, Monte-Carlo perturbation (4 cycles)"),
    4: dict(tag="configs[4]", metric="decoys_per_sec_batch_mode", desc="batch mode: 64 synthetic targets L=100..500 (name_lst), 100 decoys each",
            n_targets=64, decoys_per_target=100, seed=1000),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fold", choices=["restraint", "fold"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--decoys", type=int, default=None, help="decoys per step: per GPU (weak) or in total (strong)")
    ap.add_argument("--resident", type=int, default=None, help="positions of the resident batch per GPU")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-k1-standalone", action="store_true")
    ap.add_argument("--lbfgs-m", type=int, default=20)
    ap.add_argument("--streams", type=int, default=None, help="independent fold batches per GPU (own stream each); batch mode: targets in flight")
    ap.add_argument("--targets", type=int, default=None, help="config 4: number of targets (default 64)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per decoy evaluation of the restraint kernel, from the committed ncu --set full capture of the
    full-batch launch (profiles/k1_traffic.json, written from the .ncu-rep by tools/summarize_ncu.py)."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


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


# ------------------------------------------------------------------------------------------------ workloads
def fold_workload(cfg):
    """(seq, npzs, native) of a single-target configuration."""
    import trx2dyn  # noqa: F401
    from trx2dyn import synth
    seq, npzs, nat = synth.target(cfg["L"], cfg["seed"], dense=False, two_model=cfg["two_model"])
    if cfg["dist_only"]:
        npzs = [{"dist": z["dist"]} for z in npzs]
    return seq, npzs, nat


def fold_schedule(cfg):
    from trx2dyn import schedule
    if cfg["mc"]:
        return schedule.mc_schedule(mc_max_iter=cfg["mc"]["mc_max_iter"])
    return schedule.reference_schedule()


def batch_targets(n_targets, seed0):
    """config 4: n synthetic targets, L ~ U{100..500} (seeds 1000.., one npz each)."""
    import trx2dyn  # noqa: F401
    from trx2dyn import synth
    rng = np.random.default_rng(seed0)
    Ls = rng.integers(100, 501, size=n_targets)
    out = []
    for k, L in enumerate(Ls):
        seq, npzs, _ = synth.target(int(L), seed0 + k)
        out.append(("t%03d" % k, seq, npzs))
    return out


def build_restraint_workload(n_decoys, precision, L=300, seed=300):
    """Two table sets (two-model mixing) + decoy coordinates near the synthetic native."""
    import trx2dyn  # noqa: F401
    from trx2dyn import synth, tables
    seq, npzs, nat = synth.target(L, seed, dense=False, two_model=True)
    params = tables.load_params()
    acts = []
    for npz in npzs:
        rst = tables.gen_rst(npz, params)
        acts.append(tables.active_restraints(rst, tables.select(rst, 1, L, params)))
    rng = np.random.default_rng(seed + 7)
    sig = rng.uniform(0.3, 3.0, size=(n_decoys, 1, 1, 1))
    xyz = nat[None, :, [0, 1, 3]] + rng.normal(size=(n_decoys, L, 3, 3)) * sig
    return seq, npzs, acts, xyz.astype(np.float32 if precision == 32 else np.float64)


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_sets(npzs, L, use_orient=True):
    from oracle.tables_oracle import gen_rst_oracle, select_oracle
    from oracle import restraints_oracle as ro
    sets = []
    for npz in npzs:
        rst = gen_rst_oracle(npz, use_orient=use_orient)
        sets.append(ro.RestraintSetOracle(rst, select_oracle(rst, 1, L, 0.05), "H1"))
    return sets


def cpu_restraint_rate(npzs, xyz, n_sample, threads, L=300):
    """Oracle (CPU port of the same arithmetic) timed on a bounded sample of the workload."""
    sets = oracle_sets(npzs, L)
    half = n_sample // 2
    xs = [np.ascontiguousarray(xyz[:half], dtype=np.float64), np.ascontiguousarray(xyz[half:n_sample], dtype=np.float64)]
    sets[0].energy_grad_batch(xs[0][:threads], WEIGHTS, threads)  # warm
    t0 = time.perf_counter()
    for s, x in zip(sets, xs):
        s.energy_grad_batch(x, WEIGHTS, threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def cpu_fold_rate(npzs, seq, n_sample, threads, seed, use_orient=True):
    """Oracle fold (same schedule, fp64, one decoy per host thread) on a bounded sample."""
    from oracle import fold_oracle as fo
    sets = oracle_sets(npzs, len(seq), use_orient)
    per = [n_sample // len(sets) + (1 if k < n_sample % len(sets) else 0) for k in range(len(sets))]
    t0 = time.perf_counter()
    evals = 0
    for k, rs in enumerate(sets):
        if per[k] <= 0:
            continue
        F = fo.FoldOracle(rs, seq)
        out = F.fold(fo.random_torsions(per[k], len(seq), seed + k), fo.reference_schedule(), m=20, nthreads=threads)
        evals += int(out["evals"].sum())
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, evals


def pyrosetta_reference():
    """The reference's own implementation of the path is PyRosetta driven by folding/folding.py
    (/root/reference/folding/folding.py:48,164-171).  Usable only where both PyRosetta and a
    checkout of the reference are present (TRX_REFERENCE_DIR, default /root/reference)."""
    try:
        import pyrosetta  # noqa: F401
    except Exception:
        return None
    ref = os.environ.get("TRX_REFERENCE_DIR", "/root/reference")
    script = os.path.join(ref, "folding", "folding.py")
    return ref if os.path.exists(script) else None


def run_reference_pyrosetta(args, ref):
    """the unmodified reference folding.py (--no-fastrelax: centroid decoys, as this repo
    produces) on the example npz, one process per host core."""
    threads = os.cpu_count() or 1
    steps = args.steps or 1
    npz = os.path.join(ref, "example", "output", "seq", "pred_npz", "seq_NMR.npz")
    fasta = os.path.join(ref, "example", "seq.fasta")
    import tempfile
    t = []
    for k in range(steps):
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter()
            procs = [subprocess.Popen([sys.executable, "./folding/folding.py", "-NPZ", npz, "-FASTA", fasta, "-OUT",
                                       os.path.join(tmp, "d%d.pdb" % i), "-m", "2", "--orient", "-r", "no-idp", "--no-fastrelax"],
                                      cwd=ref, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(threads)]
            for p in procs:
                p.wait()
            t.append(time.perf_counter() - t0)
    val = threads / float(np.mean(t))
    sample = "%d decoys (one reference process per host core) of example/seq L=90, unmodified folding.py --no-fastrelax under PyRosetta" % threads
    line = {"impl": "reference", "metric": "decoys_per_sec_L90_example", "value": val, "unit": "decoys/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "example",
            "config": {"workload": "configs[0]: example/seq folding.py (PyRosetta present on this host)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "decoys/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "decoys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference_fold(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = pyrosetta_reference()
    if ref:
        return run_reference_pyrosetta(args, ref)
    cfg = CONFIGS[args.config if args.config != 4 else 2]
    threads = os.cpu_count() or 1
    steps = args.steps or 2
    n_sample = threads
    seq, npzs, _ = fold_workload(cfg)
    t = []
    for k in range(min(args.warmup, 1) + steps):
        rate, dt, _ = cpu_fold_rate(npzs, seq, n_sample, threads, cfg["seed"] + 10 * k, use_orient=not cfg["dist_only"])
        if k >= min(args.warmup, 1):
            t.append(dt)
    val = n_sample / float(np.mean(t))
    sample = ("%d decoys (one per host thread) of the %s workload per step, oracle/fold_oracle.c, same schedule%s, fp64 "
              "(PyRosetta absent: import pyrosetta fails)" % (n_sample, cfg["tag"], " without the Monte-Carlo cycles" if cfg["mc"] else ""))
    line = {"impl": "reference", "metric": cfg["metric"], "value": val, "unit": "decoys/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s)" % (cfg["desc"], cfg["tag"]), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "decoys/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "decoys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  PyRosetta is probed first; where it is
    absent (un-vendored binary dependency) this times the oracle port on all host threads."""
    if args.mode != "restraint":
        return run_reference_fold(args)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps or 3
    n_sample = max(2 * threads, 64)
    _, npzs, _, xyz = build_restraint_workload(n_sample, 64)
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


# ------------------------------------------------------------------------------------------------ GPU arm
def dist_setup():
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    return world, rank, local, barrier, max_over_ranks, sum_over_ranks


def k1_standalone(ctx_stream, L=300, seed=300, N=4096):
    """Metric M2: the restraint kernel alone on device-resident inputs, protein-like and dense tables,
    fp32 and fp64 -- decoy evaluations/s, restraint evaluations/s, algorithmic GB/s and fraction of the HBM roofline."""
    import torch
    from trx2dyn import capi, synth, tables
    peak, _ = peaks()
    out = {}
    params = tables.load_params()
    stream = torch.cuda.Stream()
    ctx = capi.Context(torch.cuda.current_device(), stream.cuda_stream)
    Lpad = capi.padded_length(L)
    G = (N + 31) // 32
    xyz = synth.random_backbones(N, L, 1)
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for dense in (False, True):
        seq, npzs, nat = synth.target(L, seed, dense=dense)
        rst = tables.gen_rst(npzs[0], params)
        tb = capi.Tables(ctx, L, tables.active_restraints(rst, tables.select(rst, 1, L, params)))
        R = sum(tb.info()["counts"])
        for prec in (32, 64):
            n = N if prec == 32 else N // 2
            g = (n + 31) // 32
            dt = torch.float32 if prec == 32 else torch.float64
            with torch.cuda.stream(stream):
                nat_d = torch.tensor(xyz[:n], dtype=dt, device="cuda")
                grp = torch.empty(g * Lpad * 9 * 32, dtype=dt, device="cuda")
                grad = torch.empty_like(grp)
                E = torch.empty(3 * g * 32, dtype=torch.float64, device="cuda")
                capi.to_grouped(ctx, n, L, 3, prec, nat_d.data_ptr(), grp.data_ptr())
                for _ in range(3):
                    tb.energy_grad_device(n, grp.data_ptr(), E.data_ptr(), grad.data_ptr(), WEIGHTS, prec)
                ctx.sync()
                ctx.set_timing(True)
                ctx.reset_timing()
                for _ in range(5):
                    flush.zero_()
                    tb.energy_grad_device(n, grp.data_ptr(), E.data_ptr(), grad.data_ptr(), WEIGHTS, prec)
                ctx.sync()
            ms, cnt = ctx.timing("restraints")
            ms2, cnt2 = ctx.timing("reduce")
            ctx.set_timing(False)
            per, per2 = ms / cnt, ms2 / cnt2
            es = prec // 8
            alg = (4 * es * R + 2 * 3 * L * 3 * es) * n
            key = "%s_f%d" % ("dense" if dense else "protein_like", prec)
            out[key] = {"decoys": n, "restraints_per_decoy": R, "kernel_ms": per, "reduce_ms": per2,
                        "decoy_evals_per_sec": n / (per + per2) * 1e3, "restraint_evals_per_sec": n * R / per * 1e3,
                        "algorithmic_GBs": alg / per / 1e6, "frac_of_hbm_peak": alg / per / 1e6 / peak}
            del nat_d, grp, grad, E
        tb.close()
    ctx.close()
    return out


def run_b200_fold(args):
    """Metric M1: fully minimised centroid decoys per second (whole schedule, on device)."""
    import torch
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    world, rank, local, barrier, max_over_ranks, sum_over_ranks = dist_setup()
    import trx2dyn  # noqa: F401
    from trx2dyn import capi, metrics, parallel, sampler, tables

    cfg = CONFIGS[args.config]
    L = cfg["L"]
    strong = args.scaling == "strong"
    if strong:
        total = args.decoys or (4096 if args.config == 2 else cfg["decoys"])
        N = parallel.shard_counts(total, world)[rank]
    else:
        N = args.decoys or cfg["decoys"]
        total = N * world
    resident = min(args.resident or cfg["resident"], N)
    steps = args.steps or 3
    warmup = max(args.warmup, 1)   # the contract asks for >= 3 (the default); fewer only for exploratory runs
    seq, npzs, nat = fold_workload(cfg)
    params = tables.load_params()
    runs = fold_schedule(cfg)
    mc = cfg["mc"]
    ntab = len(npzs)
    # `streams` independent fold batches (own context + CUDA stream each, one host thread each) share the GPU
    S = max(1, args.streams or 1)
    per = [(N // S + (1 if k < N % S else 0)) for k in range(S)]
    offs = np.concatenate([[0], np.cumsum(per)]).astype(int)

    def split(n):   # decoys of a lane over the table blocks (all but the last a multiple of 32)
        if ntab == 1:
            return [n]
        half = (n // 2 + 31) // 32 * 32
        return [half, n - half]

    lanes = []
    for k in range(S):
        stream = torch.cuda.Stream()
        ctx_k = capi.Context(local, stream.cuda_stream)
        tabs_k = [sampler.build_tables(ctx_k, npz, seq, params) for npz in npzs]
        nq_k = split(per[k])
        cap_k = split(min(resident // S if S > 1 else resident, per[k]))
        batch_k = capi.FoldBatch(ctx_k, tabs_k, cap_k, sampler.aa_index(seq), runs, lbfgs_m=args.lbfgs_m)
        lanes.append(dict(stream=stream, ctx=ctx_k, tabs=tabs_k, nq=nq_k, cap=cap_k, batch=batch_k, rounds=C.c_int(),
                          nq_arr=(C.c_int * ntab)(*nq_k)))
    R = [sum(t.info()["counts"]) for t in lanes[0]["tabs"]]
    pool_exec = ThreadPoolExecutor(max_workers=S)

    # pinned host buffers: torsions in, coordinates + terms out (the e2e path IS the product call)
    tors_h = torch.empty((N, L, 3), dtype=torch.float32).pin_memory()
    xyz_h = torch.empty((N, L, 5, 3), dtype=torch.float32).pin_memory()
    terms_h = torch.empty((N, 8), dtype=torch.float64).pin_memory()
    ncol = 3 if mc else 2
    stats_h = torch.empty((N, ncol), dtype=torch.int64).pin_memory()
    id0 = int(np.sum(parallel.shard_counts(total, world)[:rank])) if strong else rank * N

    def fold_lane(k):
        ln, o, n = lanes[k], int(offs[k]), per[k]
        a = (C.c_void_p(tors_h[o:o + n].data_ptr()), C.c_void_p(xyz_h[o:o + n].data_ptr()),
             C.c_void_p(terms_h[o:o + n].data_ptr()), C.c_void_p(stats_h[o:o + n].data_ptr()))
        if mc:
            capi.check(capi.lib().trx_fold_mc_queue(ln["batch"]._h, ln["nq_arr"], *a, C.c_int(len(runs) - 1), C.c_int(mc["cycles"]),
                                                    C.c_double(mc["kT"]), C.c_int(mc["block"][0]), C.c_int(mc["block"][1]),
                                                    C.c_double(mc["sigma_deg"]), C.c_ulonglong(cfg["seed"]), C.c_ulonglong(id0 + o),
                                                    C.c_int(1 << 30), C.c_int(16), C.byref(ln["rounds"])))
        else:
            capi.check(capi.lib().trx_fold_run_queue(ln["batch"]._h, ln["nq_arr"], *a, C.c_int(1 << 30), C.c_int(16), C.byref(ln["rounds"])))

    score_w = np.array(list(runs[-1].w))

    def one_step(seed):
        tors_h.copy_(torch.from_numpy(sampler.random_torsions(N, L, seed)))
        t0 = time.perf_counter()
        list(pool_exec.map(fold_lane, range(S)))
        if world > 1:
            # the path's only exchange: all-gather of per-decoy energies (NCCL) for pool selection
            score = terms_h.numpy() @ score_w
            full = parallel.gather_scalars(score, total, rank, world, device="cuda")
            parallel.select_pool(full[:, 0], 10)
        return time.perf_counter() - t0

    for k in range(warmup):
        one_step(1000 * rank + k)
    barrier()
    sampler_clk = ClockSampler(local)
    sampler_clk.start()
    # timed region: CUDA events only around the restraint kernel's launches (the roofline kernel) and around each fold
    # (timing mode 2); the other kernels' shares come from one more, untimed, step with every launch bracketed
    for ln in lanes:
        ln["ctx"].set_timing(2)
        ln["ctx"].reset_timing()
    launches0 = sum(ln["ctx"].launch_count for ln in lanes)
    barrier()
    t_wall, evals_total, k1_decoy_evals, rest_evals = [], 0.0, 0.0, 0.0
    for k in range(steps):
        t_wall.append(one_step(1000 * rank + 100 + k))
        evals_total += float(stats_h[:, 0].numpy().astype(np.float64).sum())
        for ln in lanes:
            ke = ln["batch"].k1_evals()
            k1_decoy_evals += float(sum(ke))
            rest_evals += float(sum(e * r for e, r in zip(ke, R)))
    barrier()
    clocks = sampler_clk.stop()
    launches = sum(ln["ctx"].launch_count for ln in lanes) - launches0
    # device time of a step: the slowest lane's fold (CUDA events on its stream, inputs resident);
    # lanes run concurrently, so steps cost max over lanes, not the sum
    t_dev = max(ln["ctx"].timing("fold_device")[0] / 1e3 for ln in lanes)
    k1_ms = sum(ln["ctx"].timing("restraints")[0] for ln in lanes)
    k1_n = sum(ln["ctx"].timing("restraints")[1] for ln in lanes)
    t_e2e = float(sum(t_wall))
    names = ("restraints", "reduce", "nerf", "centroid", "torsion_grad", "lbfgs", "cart_gather", "cart_grad", "segment",
             "compact", "activity", "turnover", "migrate")
    for ln in lanes:
        ln["ctx"].set_timing(1)
        ln["ctx"].reset_timing()
    one_step(1000 * rank + 100)   # the first timed step again, every launch bracketed: kernel shares
    busy = {name: sum(ln["ctx"].timing(name)[0] for ln in lanes) for name in names}
    counts = {name: sum(ln["ctx"].timing(name)[1] for ln in lanes) for name in names}
    tot_busy = sum(busy.values())
    shares = {name: v / tot_busy for name, v in busy.items()}
    for ln in lanes:
        ln["ctx"].set_timing(False)
    k1_share = k1_ms / (1e3 * t_dev)   # of this rank's device time of the timed steps
    t_dev, t_e2e = max_over_ranks([t_dev, t_e2e])
    evals_all, k1_all, rest_all = sum_over_ranks([evals_total, k1_decoy_evals, rest_evals])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    ca = xyz_h[: min(N, 64), :, 1].numpy().astype(np.float64)
    tm = [metrics.tm_score(c, nat[:, 1]) for c in ca[:32]]
    peak, peak_src = peaks()
    # algorithmic bytes the restraint kernel processed on THIS rank: 16 B (4 fp32 knot scalars) per restraint
    # evaluated for a decoy + coordinates in / gradient out per decoy evaluation.  Counted on the device: only the
    # evaluations the kernel actually made (vdw-only runs skip it; closing evaluations are included).
    alg_bytes = 16.0 * rest_evals + 2 * 3 * L * 12.0 * k1_decoy_evals
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tr = ncu_traffic()
    if tr and args.config == 2:
        traffic = {"bytes_per_launch": tr["dram_bytes_per_decoy_eval"] * k1_decoy_evals / max(k1_n, 1), "source": tr["source"]}
    line = {"metric": cfg["metric"], "value": total * steps / t_dev, "unit": "decoys/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, %d decoys per step %s folded through %d resident positions per GPU (continuous batching), "
                                   "full mode-2 centroid schedule incl. min_mover_cart (%s)"
                                   % (cfg["desc"], total if strong else N, "in total" if strong else "per GPU", resident, cfg["tag"]),
                       "restraints_per_decoy": R,
                       "l2": "working set per step (%.1f GB of decoy state) exceeds L2" % (batch_bytes(resident, L, args.lbfgs_m) / 1e9),
                       "streams": S, "mode": "fold", "lbfgs_m": args.lbfgs_m,
                       "monte_carlo": mc,
                       "non_restraint_terms": "vdw/rama/omega/cart_bonded/cen_hb-like backbone H-bond term: stated approximations (include/trx_centroid_model.h)",
                       "collective": "all-gather of per-decoy energies for pool selection (N>1 only)"},
            "restraint_kernel_decoy_evals_per_sec": k1_all / t_dev,
            "restraint_evals_per_sec": rest_all / t_dev,
            "mean_evals_per_decoy": evals_all / (total * steps), "rounds_last_step": lanes[0]["rounds"].value,
            "decoy_quality": {"tm_vs_synthetic_native_median": float(np.median(tm)), "tm_gt_0.5_frac": float(np.mean(np.array(tm) > 0.5))},
            "clocks": clocks,
            "e2e": {"value": total * steps / t_e2e, "unit": "decoys/s",
                    "h2d_bytes_per_step": int(tors_h.numel() * 4),
                    "d2h_bytes_per_step": int(xyz_h.numel() * 4 + tors_h.numel() * 4 + terms_h.numel() * 8 + stats_h.numel() * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "kernel": "restraints_kernel<float>", "kernel_ms": k1_ms / max(k1_n, 1),
                         "kernel_ms_total": k1_ms, "kernel_launches": k1_n,
                         "kernel_share_of_step": k1_share, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / max(k1_n, 1),
                         "accounting": "achieved = (16 B x restraint evaluations + 7200 B x decoy evaluations the kernel made, counted on the "
                                       "device by compact_kernel, rank 0) / summed CUDA-event time of the kernel's launches (rank 0)",
                         "kernel_decoy_evals": k1_decoy_evals, "kernel_restraint_evals": rest_evals,
                         "kernel_shares": shares, "kernel_ms_by_name": {k: round(v, 2) for k, v in busy.items()},
                         "kernel_launches_by_name": counts,
                         "kernel_shares_note": "shares / ms_by_name / launches_by_name: one extra untimed step with every launch bracketed by events"}}
    if world == 1 and not args.no_k1_standalone and args.config == 2:
        line["k1_standalone"] = k1_standalone(None)
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only (the other ranks of an N>1 run would wait on it)
        threads = os.cpu_count() or 1
        rate, dt, ev = cpu_fold_rate(npzs, seq, threads, threads, cfg["seed"], use_orient=not cfg["dist_only"])
        line["cpu_baseline"] = {"value": rate, "unit": "decoys/s", "cores": threads, "kind": "port",
                                "sample": "%d decoys (one per host thread) of the same workload%s, oracle/fold_oracle.c same schedule fp64, %.1f s (PyRosetta absent)"
                                          % (threads, " without the Monte-Carlo cycles" if mc else "", dt)}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_b200_batch_mode(args):
    """configs[4]: name_lst batch mode -- many targets of different length, 100 decoys each, target-and-decoy sharded
    over the ranks (parallel.assign_blocks), several targets in flight per GPU (one CUDA stream each)."""
    import torch
    world, rank, local, barrier, max_over_ranks, sum_over_ranks = dist_setup()
    import trx2dyn  # noqa: F401
    from trx2dyn import capi, pipeline
    cfg = CONFIGS[4]
    n_targets = args.targets or cfg["n_targets"]
    nd = args.decoys or cfg["decoys_per_target"]
    targets = batch_targets(n_targets, cfg["seed"])
    n_dec = [nd] * n_targets
    S = max(1, args.streams or 4)
    streams = [torch.cuda.Stream() for _ in range(S)]
    ctxs = [capi.Context(local, st.cuda_stream) for st in streams]
    steps = args.steps or 1
    warmup = max(args.warmup, 1)

    def one_step(seed):
        t0 = time.perf_counter()
        res = pipeline.fold_batch(ctxs, targets, n_dec, rank, world, seed=seed)
        return time.perf_counter() - t0, res

    for k in range(warmup):
        one_step(7 + k)
    barrier()
    clk = ClockSampler(local)
    clk.start()
    launches0 = sum(c.launch_count for c in ctxs)
    for c in ctxs:
        c.set_timing(True)
        c.reset_timing()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall, mine = [], 0
    for k in range(steps):
        torch.cuda.synchronize()
        ev[k][0].record()
        dt, res = one_step(100 + k)
        torch.cuda.synchronize()
        ev[k][1].record()
        t_wall.append(dt)
        mine = len(res)
    barrier()
    clocks = clk.stop()
    launches = sum(c.launch_count for c in ctxs) - launches0
    k1_ms = sum(c.timing("restraints")[0] for c in ctxs)
    dev_busy = sum(c.timing("fold_device")[0] for c in ctxs) / 1e3
    for c in ctxs:
        c.set_timing(False)
    torch.cuda.synchronize()
    t_evt = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    t_evt, t_e2e = max_over_ranks([t_evt, float(sum(t_wall))])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    total = sum(n_dec)
    Ls = [len(t[1]) for t in targets]
    line = {"metric": cfg["metric"], "value": total * steps / t_evt, "unit": "decoys/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * t_evt / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s; %d targets x %d decoys, L in [%d, %d] (mean %.0f), target-and-decoy sharded over %d GPU(s), %d targets in flight per GPU (%s)"
                                   % (cfg["desc"], n_targets, nd, min(Ls), max(Ls), float(np.mean(Ls)), world, S, cfg["tag"]),
                       "timing": "a step is host-driven (tables of 64 targets are built and uploaded inside it): value is timed with CUDA events around the step on the default stream, inputs (npz) on the host",
                       "rank0_decoys": mine, "device_busy_s_sum_over_streams": dev_busy, "restraint_kernel_ms": k1_ms},
            "clocks": clocks,
            "e2e": {"value": total * steps / t_e2e, "unit": "decoys/s",
                    "h2d_bytes_per_step": int(sum(n * L * 12 for n, L in zip(n_dec, Ls)) // world),
                    "d2h_bytes_per_step": int(sum(n * L * (60 + 12) + n * 72 for n, L in zip(n_dec, Ls)) // world)},
            "gpu_launches": int(launches), "roofline": None}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def batch_bytes(N, L, m):
    return N * L * 15 * 4.0 * (5 + 2 * m) + N * L * 15 * 4.0 * 3


def run_b200(args):
    if args.mode != "restraint":
        return run_b200_batch_mode(args) if args.config == 4 else run_b200_fold(args)
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

    L_TARGET = 300
    N, prec = args.decoys or 4096, args.precision
    steps = args.steps or 10
    seq, npzs, acts, xyz = build_restraint_workload(N, prec)
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
