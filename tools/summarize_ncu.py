"""Turns ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

  python tools/summarize_ncu.py full gpurun_out/k1_r1h_sparse.ncu-rep "header line"   > profiles/....txt
  python tools/summarize_ncu.py launches gpurun_out/launches_r1h.csv "header line"    > profiles/....md
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, header):
    print("#", header)
    print("# source:", rep, "(ncu --set full --clock-control none --import-source on; clocks not locked)\n")
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    print("kernel:", d["Kernel Name"][0])
    for k in WANT:
        if k in d:
            print("%-70s %s %s" % (k, d[k][0], d[k][1]))
    st = {}
    for h, (v, _) in d.items():
        if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            st[h.split("issue_stalled_")[1].split("_per_issue")[0]] = float(v)
    print("\nstall reasons (warps per issue-active cycle):")
    for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
        print("  %-22s %.3f" % (k, v))
    rows = ncu_csv(rep, "source")
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ie, src, ns = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    mix, tot, samples = collections.Counter(), 0, []
    for r in rows[hi + 1:]:
        if len(r) <= ie:
            continue
        toks = r[src].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        n = int(r[ie])
        mix[op] += n
        tot += n
        samples.append((int(r[ns]), r[src].strip()))
    print("\ninstruction mix (warp-level executed, %d total):" % tot, ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in mix.most_common(14)))
    ssum = sum(s for s, _ in samples)
    print("\ntop SASS lines by stall samples (of %d):" % ssum)
    for s, line in sorted(samples, key=lambda x: -x[0])[:12]:
        print("  %5.2f%%  %s" % (100.0 * s / max(ssum, 1), line[:70]))


def launches(path, header):
    print("#", header, "\n")
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric, Unit, Value
    agg = collections.OrderedDict()
    lst = []
    for r in rows:
        name, val, unit = r[4], float(r[-1].replace(",", "")), r[-2]
        us = val / 1e3 if unit in ("nsecond", "ns") else (val * 1e3 if unit in ("msecond", "ms") else val)
        name = name.replace("void trx::", "").split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        lst.append((r[0], name, r[8], r[7], us))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.1f | %.3f | %.1f |" % (k, n, us, us / tot, us / n))
    print("\nFirst 40 launches of the window (ID, kernel, grid, block, us):\n\n```")
    for r in lst[:40]:
        print("%s  %s  %s  %s  %.2f" % r)
    print("```")


def traffic(rep, decoys):
    """profiles/k1_traffic.json: DRAM bytes per decoy evaluation of the restraint kernel, from one ncu --set full capture of
    the full-batch standalone launch (bench.py multiplies it by the decoy evaluations per launch of the fold)."""
    import json
    rows = ncu_csv(rep, "raw")
    d = dict(zip(rows[0], zip(rows[2], rows[1])))

    def to_bytes(key):
        v, unit = d[key]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return float(v.replace(",", "")) * mult
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    print(json.dumps({"dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "decoys_per_launch": int(decoys),
                      "dram_bytes_per_decoy_eval": (rd + wr) / float(decoys), "kernel": d["Kernel Name"][0],
                      "source": "ncu --set full capture %s (tools/k1_bench.py, L=300 protein-like table, %s decoys): dram__bytes_read.sum + dram__bytes_write.sum" % (rep, decoys)}, indent=1))


if __name__ == "__main__":
    {"full": full, "launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
