"""Per-CUDA-source-line instruction counts and stall samples of one kernel in an ncu report (needs -lineinfo and
--import-source on).  usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP] [LAUNCH_INDEX]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern] + (["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []),
                     capture_output=True, text=True).stdout
agg = collections.OrderedDict()
hdr = None
fname = ""
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {c: i for i, c in enumerate(r)}
        hdr["Source"] = 1
    elif hdr and len(r) > 8 and r[2] == "-":               # line-level aggregate row
        key = (fname, r[0], r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[hdr["Instructions Executed"]] or 0)
        a[1] += int(r[hdr["# Samples"]] or 0)
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instructions %d, samples %d" % (ti, ts))
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * n / max(ti, 1), 100.0 * s / max(ts, 1), f, ln, src))
