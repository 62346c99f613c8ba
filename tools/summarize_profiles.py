"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01a      # tag = round / visit

Writes  profiles/<tag>_launches.csv        every launch of our kernels in one bench step (name, grid, duration us)
        profiles/<tag>_launches_summary.md share of the step per kernel, per-node table
        profiles/<tag>_ncu_full.csv        selected metrics of the `ncu --set full` captures (incl. DRAM bytes)
        profiles/ncu_dominant_kernel.json  DRAM bytes per launch of the dominant kernel (bench.py roofline.traffic)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", name))[:80]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    import b200quant.workloads as w
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[short(r["Kernel Name"]).split("<")[0]][0] += 1
        agg[short(r["Kernel Name"]).split("<")[0]][1] += float(r["Metric Value"].replace(",", "")) / 1e3
    ours = [r for r in rows if re.search(r"(flat|seg|ew_kernel|multi|threshold_update|foldbn|wnq|qil)", r["Kernel Name"])
            and "at::" not in r["Kernel Name"]]
    nodes = w.resnet50_nodes(256)
    per_step = 3 * len(nodes)
    nsteps = len(ours) // per_step
    step = ours[(nsteps - 1) * per_step: nsteps * per_step] if nsteps else ours
    with open(os.path.join(PROF, tag + "_launches.csv"), "w") as f:
        f.write("idx,kernel,grid,block,duration_us\n")
        for i, r in enumerate(step):
            f.write("%d,%s,%s,%s,%.3f\n" % (i, short(r["Kernel Name"]), r["Grid Size"].replace(",", " "),
                                            r["Block Size"].replace(",", " "), float(r["Metric Value"].replace(",", "")) / 1e3))
    tot = sum(v[1] for k, v in agg.items() if "at::" not in k)
    md = ["# %s: ncu launch list of `bench.py --steps 2 --warmup 3 --profile`" % tag, "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: shares, not absolutes).",
          "", "| kernel | launches | total us | share of our kernels | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "at::" in k:
            continue
        md.append("| %s | %d | %.1f | %.1f%% | %.2f |" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    if nsteps and len(step) == per_step:
        md += ["", "## last captured step, per node (forward = reduction + QDQ sweep, backward = STE copy)", "",
               "| node | shape | elements | reduce us | GB/s (4 B/el) | qdq us | GB/s (8 B/el) | bwd us | GB/s (8 B/el) |",
               "|---|---|---|---|---|---|---|---|---|"]
        fwd, bwd = step[:2 * len(nodes)], step[2 * len(nodes):]
        tr = tq = tb = 0.0
        for i, (name, kind, shape) in enumerate(nodes):
            n = w.numel(shape)
            r = float(fwd[2 * i]["Metric Value"].replace(",", "")) / 1e3
            q = float(fwd[2 * i + 1]["Metric Value"].replace(",", "")) / 1e3
            b = float(bwd[len(nodes) - 1 - i]["Metric Value"].replace(",", "")) / 1e3
            tr, tq, tb = tr + r, tq + q, tb + b
            md.append("| %s | %s | %d | %.2f | %.0f | %.2f | %.0f | %.2f | %.0f |" %
                      (name, "x".join(map(str, shape)), n, r, 4 * n / r / 1e3, q, 8 * n / q / 1e3, b, 8 * n / b / 1e3))
        md.append("| **total** | | | **%.1f** | | **%.1f** | | **%.1f** | |" % (tr, tq, tb))
    open(os.path.join(PROF, tag + "_launches_summary.md"), "w").write("\n".join(md) + "\n")


def full(tag):
    out_rows = []
    dom = None
    for rep in sorted(os.listdir(OUT)):
        if not rep.endswith(".ncu-rep"):
            continue
        res = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True,
                             text=True)
        rows = list(csv.reader(res.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"report": rep, "kernel": short(r[hdr.index("Kernel Name")])}
            for m in METRICS:
                if m in hdr:
                    rec[m] = r[hdr.index(m)]
                    rec[m + "_unit"] = units[hdr.index(m)]
            out_rows.append(rec)
            if "qdq_flat_hot" in rec["kernel"] and "dram__bytes_read.sum" in rec:
                def to_bytes(v, u):
                    v = float(v.replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                rd = to_bytes(rec["dram__bytes_read.sum"], rec["dram__bytes_read.sum_unit"])
                wr = to_bytes(rec["dram__bytes_write.sum"], rec["dram__bytes_write.sum_unit"])
                tot = rd + wr
                if dom is None or tot > dom["dram_bytes_per_launch"]:
                    dom = {"kernel": rec["kernel"], "dram_bytes_per_launch": tot, "dram_bytes_read": rd,
                           "dram_bytes_write": wr,
                           "alg_bytes_of_this_launch": 2.0 * rd,   # the sweep reads 4 B and writes 4 B per element
                           "comment": "writes below the algorithmic 4 B/element: part of the output is still dirty in "
                                      "the 126 MB L2 when the kernel ends and is written back later",
                           "duration_us": float(rec["gpu__time_duration.sum"].replace(",", "")), "source": tag + "/" + rep,
                           "note": "largest captured launch of the dominant kernel; ncu --set full --clock-control none"}
    if out_rows:
        keys = ["report", "kernel"] + [k for m in METRICS for k in (m, m + "_unit")]
        with open(os.path.join(PROF, tag + "_ncu_full.csv"), "w") as f:
            wtr = csv.DictWriter(f, fieldnames=keys)
            wtr.writeheader()
            for r in out_rows:
                wtr.writerow({k: r.get(k, "") for k in keys})
    if dom:
        json.dump(dom, open(os.path.join(PROF, "ncu_dominant_kernel.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag)
    # only what the visit being summarised produced: gpurun_out/ is scratch and keeps files of earlier visits
    ref = os.path.join(OUT, "launches.csv")
    newest = os.path.getmtime(ref) if os.path.exists(ref) else 0.0
    for f in ("bench_ours.json", "bench_reference.json", "microbench.log", "l2_probe.json"):
        p = os.path.join(OUT, f)
        if os.path.exists(p) and os.path.getsize(p) and abs(os.path.getmtime(p) - newest) < 3600:
            open(os.path.join(PROF, tag + "_" + f), "w").write(open(p).read())
    print("profiles/: " + ", ".join(sorted(os.listdir(PROF))))
