"""Markdown table of the key metrics of EVERY launch in an ncu report (raw page), read here without a GPU.
usage: ncu_table.py REPORT.ncu-rep ["title"]            (prints markdown)
       ncu_table.py REPORT.ncu-rep --traffic KEY        (adds {KEY: dram bytes per step = sum over the report's launches}
                                                         to profiles/traffic.json; with --kernel REGEX only those launches,
                                                         --per-launch divides by their number)"""
import csv, io, json, os, re, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "": 1.0, "%": 1.0}
COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("launch__registers_per_thread", "regs")]


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    out = []
    for vals in r[2:]:
        if len(vals) != len(hdr):
            continue
        out.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
    return out


def num(cell):
    v, u = cell
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(u, 1.0)


def main():
    rep = sys.argv[1]
    rows = rows_of(rep)
    if "--traffic" in sys.argv:
        key = sys.argv[sys.argv.index("--traffic") + 1]
        pat = re.compile(sys.argv[sys.argv.index("--kernel") + 1]) if "--kernel" in sys.argv else None
        sel = [m for m in rows if (pat is None or pat.search(m["Kernel Name"][0])) and num(m["dram__bytes_read.sum"]) is not None]
        tot = sum(num(m["dram__bytes_read.sum"]) + num(m["dram__bytes_write.sum"]) for m in sel)
        if "--per-launch" in sys.argv:
            tot /= max(len(sel), 1)
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        t = json.load(open(path)) if os.path.exists(path) else {}
        t[key] = tot
        t.setdefault("_source", {})[key] = "%s: %d launch(es)%s" % (os.path.basename(rep), len(sel),
                                                                    " matching " + pat.pattern if pat else "")
        json.dump(t, open(path, "w"), indent=1, sort_keys=True)
        print(key, tot)
        return
    title = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
    print("## %s\n\nsource: `%s` (ncu --set full --clock-control none; per-launch times are cold-cache and serialised)\n"
          % (title, os.path.basename(rep)))
    print("| # | kernel | grid | block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|---|---|" + "---|" * len(COLS))
    for i, m in enumerate(rows):
        if num(m.get("gpu__time_duration.sum", ("", ""))) is None:
            continue
        cells = []
        for k, label in COLS:
            x = num(m[k]) if k in m else None
            if x is None:
                cells.append("")
            elif "MB" in label:
                cells.append("%.1f" % (x / 1e6))
            elif label == "ms":
                cells.append("%.4f" % x)
            else:
                cells.append("%.1f" % x)
        name = m["Kernel Name"][0].split("(")[0]
        print("| %d | `%s` | %s | %s | %s |" % (i, name, m["Grid Size"][0], m["Block Size"][0], " | ".join(cells)))
    print()


main()
