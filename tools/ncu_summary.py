"""Summarise ncu captures brought back in gpurun_out/ into profiles/ (tracked).

  python tools/ncu_summary.py <tag> [--rep gpurun_out/prof.ncu-rep] [--launches gpurun_out/launches.csv]

Writes profiles/<tag>_kernels.csv (one line per captured launch with the metrics the roofline
is argued from), profiles/<tag>_launches.txt (per-kernel share of the launch list) and refreshes
profiles/ncu_summary.json (per-kernel DRAM bytes per launch, read by bench.py for roofline.traffic).
"""
import argparse, csv, io, json, os, subprocess, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]

VARIANTS = {(0, 8, 1): "loop-unroll4/8cta", (2, 8, 1): "chunk2/8cta", (4, 6, 1): "chunk4/6cta", (4, 4, 1): "chunk4/4cta",
            (8, 4, 1): "chunk8/4cta", (4, 3, 2): "rows2-chunk4/3cta", (2, 4, 2): "rows2-chunk2/4cta"}


STRIDE = {(2, 8, 1): "stride-chunk2/8cta", (4, 6, 1): "stride-chunk4/6cta"}


def short(name):
    """k_primal<kWriteD, kDict, kChunk, kMinB, kComm, kRows> / k_dual<kDict, kChunk, kMinB, kComm, kRows> ->
    'k_primal[chunk2/8cta]' (+ ',dict' / ',write_d' / ',fused-halo'): the key bench.py looks its traffic up with."""
    import re

    m = re.search(r"(k_primal|k_dual)<([^>]*)>", name)
    if m:
        t = [int(v.replace("(bool)", "").replace("(int)", "")) for v in m.group(2).split(",")]
        if m.group(1) == "k_primal":
            t += {5: [1, 0], 6: [0]}.get(len(t), [])  # (kRows and kPersist have defaults: older captures lack them)
            write_d, dict_, chunk, minb, comm, rows, persist = t[:7]
        else:
            write_d = 0
            t += {4: [1, 0], 5: [0]}.get(len(t), [])
            dict_, chunk, minb, comm, rows, persist = t[:6]
        table = STRIDE if persist else VARIANTS
        tags = [table.get((chunk, minb, rows), "chunk%d/%dcta/rows%d" % (chunk, minb, rows))]
        tags += ["dict"] * bool(dict_) + ["write_d"] * bool(write_d) + ["fused-halo"] * bool(comm)
        return "%s[%s]" % (m.group(1), ",".join(tags))
    m = re.search(r"(k_primal_band|k_dual_band)<([^>]*)>", name)
    if m:
        return "%s<%s>" % (m.group(1), m.group(2).replace(" ", "").replace("(bool)", ""))
    for k in ("k_stats_rows", "k_stats_cols", "k_stats_final", "k_precond", "k_fill_sell", "k_push", "k_wait", "k_long_partial",
              "k_long_finish", "k_tiny_iterate", "k_cluster_iterate"):
        if k in name:
            return k
    return name.split("(")[0][-60:]

def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("tbyte", 1e12), ("byte", 1.0)):
        if u.startswith(pre):
            return v * f
    return v

ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--rep", default=None)
ap.add_argument("--launches", default=None)
a = ap.parse_args()
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
summary_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
summary = json.load(open(summary_path)) if os.path.isfile(summary_path) else {"kernels": {}}

if a.rep:
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out_rows, per, extra = [], collections.defaultdict(list), {}
    for r in body:
        name = short(r[idx["Kernel Name"]])
        rec = {"kernel": name}
        for m in METRICS:
            if m in idx:
                rec[m] = r[idx[m]] + " " + units[idx[m]]
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        rec["dram_bytes_total"] = rd + wr
        per[name].append(rd + wr)
        extra.setdefault(name, []).append((float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]],
                                           float(r[idx["lts__t_sector_hit_rate.pct"]].replace(",", ""))))
        out_rows.append(rec)
    with open(os.path.join(ROOT, "profiles", a.tag + "_kernels.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["kernel", "dram_bytes_total"] + [m for m in METRICS if m in idx])
        w.writeheader()
        for rec in out_rows:
            w.writerow(rec)
    for name, vals in per.items():
        summary["kernels"][name] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
                                    "duration_%s" % extra[name][0][1]: sum(e[0] for e in extra[name]) / len(vals),
                                    "l2_sector_hit_rate_pct": sum(e[2] for e in extra[name]) / len(vals),
                                    "source": a.tag + "_kernels.csv (ncu --set full --clock-control none)"}
    print("captured launches:", {k: len(v) for k, v in per.items()})

if a.launches:
    lines = [l for l in open(a.launches) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)
        k = short(r[idx["Kernel Name"]])
        tot[k][0] += 1
        tot[k][1] += v
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(ROOT, "profiles", a.tag + "_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("%-40s %8s %14s %8s %12s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        for k, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s %8d %14.1f %7.1f%% %12.1f\n" % (k, cnt, us, 100 * us / total, us / cnt))
    print(open(os.path.join(ROOT, "profiles", a.tag + "_launches.txt")).read())

json.dump(summary, open(summary_path, "w"), indent=1)
