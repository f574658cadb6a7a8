"""Summarise EVERY launch of an ncu report (read here, no GPU needed) as one markdown table, and optionally record the
summed DRAM traffic in profiles/traffic.json (the number bench.py prints as roofline.traffic).
usage: ncu_multi.py REPORT.ncu-rep "title" [--traffic KEY] [--skip N] [--count N] >> profiles/xxx.md"""
import csv, io, json, os, subprocess, sys

rep, title = sys.argv[1], sys.argv[2]
opt = sys.argv[3:]
key = opt[opt.index("--traffic") + 1] if "--traffic" in opt else None
skip = int(opt[opt.index("--skip") + 1]) if "--skip" in opt else 0
count = int(opt[opt.index("--count") + 1]) if "--count" in opt else 10 ** 9
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def scale(k):                      # bytes / time units differ per report: normalise to MB and us
    u = units[ix[k]].lower() if k in ix else ""
    return {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
            "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)


cols = [("time us", "gpu__time_duration.sum"), ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM written MB", "dram__bytes_write.sum"),
        ("regs", "launch__registers_per_thread"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fp64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("shared wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]
cols = [c for c in cols if c[1] in ix]
print("## %s\n" % title)
print("source: `%s` (ncu --set full --clock-control none --import-source on; per-launch times under ncu are cold-cache and "
      "serialised)\n" % os.path.basename(rep))
print("| kernel | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
tot_t = tot_r = tot_w = 0.0
for r in rows[2 + skip: 2 + skip + count]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("pdsb::", "")
    vals = []
    for label, k in cols:
        v = num(r, k)
        if k.startswith("gpu__time") or k.startswith("dram__bytes"):
            v *= scale(k)
        vals.append("%.1f" % v if v == v else "-")
    tot_t += num(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum")
    tot_r += num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    tot_w += num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    print("| `%s` | " % name + " | ".join(vals) + " |")
print("\nsum over the launches: %.1f us, DRAM read %.1f MB + written %.1f MB = %.1f MB\n" % (tot_t, tot_r, tot_w, tot_r + tot_w))
st_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")]
print("warp stall reasons per issue-active cycle, launches above 20 us:\n")
for r in rows[2 + skip: 2 + skip + count]:
    if num(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum") < 20:
        continue
    st = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), num(r, h))
                 for h in st_keys), key=lambda kv: -kv[1])
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("pdsb::", "")
    print("* `%s`: " % name + ", ".join("%s %.2f" % kv for kv in st[:6] if kv[1] >= 0.05))
print()
if key:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    t[key] = (tot_r + tot_w) * 1e6
    t.setdefault("_source", {})[key] = "%s: %d launch(es) summed" % (os.path.basename(rep), len(rows[2 + skip: 2 + skip + count]))
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
