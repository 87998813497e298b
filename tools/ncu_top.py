"""Summarise an `ncu --page source --csv` export of ONE kernel: stall-reason totals and the top SASS lines by samples.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<name> --launch-count 1 > src.csv
    python tools/ncu_top.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for i, r in enumerate(rows[:5]):
    if "Source" in r:
        h, start = r, i + 1
        break
ci = {c: h.index(c) for c in h}
body = [r for r in rows[start:] if len(r) >= len(h)]
body = body[len(body) // 2:]   # the export lists the function twice (source view, SASS view); keep one
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


body = [r for r in body if r[ci["Source"]] != "Source"]
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(num(r[ci[c]]) for r in body) for c in reasons}
allsamp = sum(num(r[ci["# Samples"]]) for r in body)
print("samples", allsamp)
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-24s %8.0f %5.1f%%" % (c, v, 100 * v / max(allsamp, 1)))
items = sorted(body, key=lambda r: -num(r[ci["# Samples"]]))[:top]
for r in items:
    rs = sorted(((num(r[ci[c]]), c[6:]) for c in reasons), reverse=True)[:2]
    print("%7.0f %5.1f%%  %-60s %s" % (num(r[ci["# Samples"]]), 100 * num(r[ci["# Samples"]]) / max(allsamp, 1),
                                      r[ci["Source"]][:60], " ".join("%s=%.0f" % (n, v) for v, n in rs)))
