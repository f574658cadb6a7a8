"""Summarise an ncu report (read here, no GPU needed) into markdown for profiles/.
usage: summarize_ncu.py REPORT.ncu-rep "title" >> profiles/xxx.md"""
import csv, io, subprocess, sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
print("## %s\n" % title)
print("source: `%s` (ncu --set full --clock-control none --import-source on)\n" % rep.split("/")[-1])
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in m:
        print("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
print("\nwarp stall reasons per issue-active cycle (smsp__average_warps_issue_stalled_*_per_issue_active):\n")
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v))
      for h, (v, u) in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")]
print(", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda kv: -kv[1]) if kv[1] >= 0.01))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
agg = {}
for r in rows[2:]:
    toks = r[1].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    a = agg.setdefault(op, [0, 0])
    a[0] += int(r[ix["# Samples"]])
    a[1] += int(r[ix["Instructions Executed"]])
tot = sum(a[0] for a in agg.values())
print("\ninstruction mix (SASS, from the source page; FFMA2 = packed fp32x2 FMA, UBLKCP = bulk TMA copy):\n")
print("| opcode | warp instructions executed | share of stall samples |\n|---|---|---|")
for k, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("| %s | %d | %.1f %% |" % (k, n, 100.0 * s / max(tot, 1)))
tma = [k for k in agg if k.startswith("UBLKCP") or k.startswith("SYNCS")]
print("\nTMA / mbarrier opcodes present: %s\n" % (", ".join(sorted(tma)) or "none"))
