"""Per-kernel share of a step from an ncu launch list (gpu__time_duration.sum per launch)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    if len(r) != len(hdr) or r[0] == "ID":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.2f %% |" % (k[:90], n, t, 100 * t / tot))
