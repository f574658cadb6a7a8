"""Top stalled SASS instructions of a kernel from an ncu report (source page).
usage: ncu_hot.py REPORT.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(out))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
col = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", tot)
agg = {k: sum(int(r[col[k]] or 0) for r in body) for k in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
body.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for r in body[:n]:
    why = {k[6:]: int(r[col[k]] or 0) for k in stalls if int(r[col[k]] or 0)}
    print(r[col["Address"]][-5:], "%6s" % r[col["# Samples"]], r[col["Source"]][:70].ljust(70), dict(sorted(why.items(), key=lambda kv: -kv[1])[:3]))
