"""Copies the JSON lines of the round's bench runs (gpurun_out/r2_*.log) into profiles/ and writes profiles/r2_bench_lines.md."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
runs = [("r2_bench_default", "python bench.py --decoys 32768  (configs[2], 32768 decoys per step through 4096 resident positions; the default is 24576)"),
        ("r2_bench_default24k", "python bench.py --resident 4096  (configs[2], 24576 decoys per step through 4096 resident positions: the default until the sweep below)"),
        ("r2_bench_default12k", "python bench.py  (DEFAULT, final build: configs[2], 24576 decoys per step through 12288 resident positions)"),
        ("r2_bench_r8192", "python bench.py --resident 8192 --steps 2 --warmup 1 --no-k1-standalone"),
        ("r2_bench_r12288", "python bench.py --resident 12288 --steps 2 --warmup 1 --no-k1-standalone"),
        ("r2_bench_r16384", "python bench.py --resident 16384 --steps 2 --warmup 1 --no-k1-standalone"),
        ("r2_bench_s2_final", "python bench.py --resident 4096 --streams 2 --steps 2 --warmup 1 --no-k1-standalone  (two lanes of 2048 positions; K1 launches of the two streams overlap, so its event-timed duration and roofline.frac are not those of a kernel running alone)"),
        ("r2_bench_s2_12k", "python bench.py --streams 2 --steps 2 --warmup 1 --no-k1-standalone --no-cpu-baseline  (final build, default workload as two lanes of 6144 positions on two streams; the lanes' K1 launches overlap, so roofline.frac is not that of a kernel alone)"),
        ("r2_bench_65536", "python bench.py --decoys 65536 --steps 1 --warmup 1  (queue of 16 resident batches)"),
        ("r2_bench_s2", "python bench.py --streams 2 --steps 1 --warmup 1  (two fold lanes on two streams)"),
        ("r2_bench_4096", "python bench.py --decoys 4096 --resident 4096  (configs[2], one resident batch: no refill)"),
        ("r2_bench_reference", "python bench.py --impl reference"),
        ("r2_bench_512", "python bench.py --decoys 512 --resident 512  (configs[2], the per-GPU share of the strong-scaling run at 8 GPUs)"),
        ("r2_bench_c1", "python bench.py --config 1  (configs[1]: L=150 distance-only, 256 decoys)"),
        ("r2_bench_c3", "python bench.py --config 3  (configs[3]: L=800, 2048 decoys, 4 MC cycles, 1 GPU)"),
        ("r2_bench_c4", "python bench.py --config 4 --streams 8  (configs[4]: 64 targets x 100 decoys, 1 GPU)"),
        ("r2_bench_c4_pool", "python bench.py --config 4 --streams 8 --steps 2 --warmup 1  (configs[4], final build: the 64 targets' tables and fold batches come from the contexts' block pools)"),
        ("r2_2gpu_c2_weak_final", "torchrun x2 bench.py --gpus 2 --steps 1 --warmup 1 --no-k1-standalone  (final build, default workload on 2 GPUs: weak scaling)"),
        ("r2_8gpu_c2_weak", "torchrun x8 bench.py --gpus 8 --decoys 16384  (configs[2], weak: 16384 decoys per GPU and step; the build before the default became 32768)"),
        ("r2_8gpu_c2_strong", "torchrun x8 bench.py --gpus 8 --scaling strong --decoys 4096  (configs[2] as written: 4096 decoys sharded over 8 GPUs)"),
        ("r2_4gpu_c2_strong", "torchrun x4 ... --scaling strong --decoys 4096"),
        ("r2_2gpu_c2_strong", "torchrun x2 ... --scaling strong --decoys 4096"),
        ("r2_8gpu_c3", "torchrun x8 bench.py --gpus 8 --config 3 --scaling strong  (configs[3]: 2048 decoys sharded over 8 GPUs)"),
        ("r2_8gpu_c4", "torchrun x8 bench.py --gpus 8 --config 4 --streams 8  (configs[4]: 64 targets x 100 decoys, target-and-decoy sharded)")]
out = ["# Round 2 bench lines (B200, gpurun; one JSON line per run, copied verbatim under profiles/r2_line_*.json)\n",
       "| run | command | value | e2e | ms/step | steps | roofline.frac (K1) | notes |", "|---|---|---|---|---|---|---|---|"]
for tag, cmd in runs:
    p = os.path.join(ROOT, "gpurun_out", tag + ".log")
    kept = os.path.join(ROOT, "profiles", "r2_line_%s.json" % tag.replace("r2_", ""))
    line = None
    if os.path.exists(p):
        for ln in open(p):
            ln = ln.strip()
            if ln.startswith("{") and '"metric"' in ln:
                line = ln
    if line is None and not os.path.exists(kept):
        continue
    d = json.loads(line) if line else json.load(open(kept))   # a line whose log is gone stays as committed
    json.dump(d, open(kept, "w"), indent=1)
    rf = d.get("roofline") or {}
    notes = []
    if "mean_evals_per_decoy" in d:
        notes.append("%.0f evals/decoy" % d["mean_evals_per_decoy"])
    if d.get("decoy_quality"):
        notes.append("median TM %.2f" % d["decoy_quality"]["tm_vs_synthetic_native_median"])
    if d.get("clocks", {}).get("sm_mhz"):
        notes.append("SM %d MHz %s" % (d["clocks"]["sm_mhz"], ",".join(d["clocks"].get("reasons", [])) or "no throttle"))
    if "cpu_baseline" in d:
        notes.append("CPU %s: %.2f %s on %d cores" % (d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], d["cpu_baseline"]["cores"]))
    out.append("| %s | `%s` | %.1f %s | %.1f | %.0f | %d | %s | %s |" % (
        tag, cmd, d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["steps"],
        ("%.3f" % rf["frac"]) if rf.get("frac") else "-", "; ".join(notes)))
open(os.path.join(ROOT, "profiles", "r2_bench_lines.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
