"""profiles/<name>.md from an `ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv` launch list: launches, total time, share,
average, DRAM bytes per launch for every kernel (cold-cache, serialised: compare SHARES, not absolutes).
usage: python scripts/launch_list.py profiles/r2_launches_raw.csv r2_launches "title" """
import collections
import csv
import re
import sys

src, name, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 6]
H = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        k = re.sub(r"\(.*", "", r[H["Kernel Name"]]).split("::")[-1]
        m, v, u = r[H["Metric Name"]], float(r[H["Metric Value"]].replace(",", "")), r[H["Metric Unit"]]
    except Exception:
        continue
    a = agg.setdefault(k, [0, 0.0, 0.0])
    if m == "gpu__time_duration.sum":
        a[0] += 1
        a[1] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    elif m.startswith("dram__bytes"):
        a[2] += v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
tot = sum(a[1] for a in agg.values())
out = ["# %s" % title, "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold-cache, serialised: compare shares)",
       "| kernel | launches | total ms | share | avg us | DRAM MB/launch |", "|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.3f | %.1f%% | %.1f | %.1f |" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0], a[2] / a[0]))
open("profiles/%s.md" % name, "w").write("\n".join(out) + "\n")
print("\n".join(out))
