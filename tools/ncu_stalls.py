"""scratch: per-instruction top stalls with reason from ncu source csv (second half = SASS view)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
body = rows[2:]
half = len(body) // 2
body = body[half:]   # the export lists the function twice; keep one copy
col = {c: h.index(c) for c in h}
reasons = ['stall_barrier','stall_long_sb','stall_short_sb','stall_wait','stall_lg','stall_mio','stall_math','stall_membar','stall_sleep','stall_branch_resolving','stall_not_selected','stall_selected','stall_dispatch','stall_drain','stall_no_inst','stall_misc','stall_tex']
tot = {r: 0 for r in reasons}
lines = []
for i, r in enumerate(body):
    try:
        s = int(r[col['# Samples']])
    except Exception:
        continue
    rs = {k: int(r[col[k]] or 0) for k in reasons}
    for k in reasons: tot[k] += rs[k]
    lines.append((s, i, r[col['Source']][:90], int(r[col['Instructions Executed']] or 0), rs))
print('totals:', {k: v for k, v in tot.items() if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for s, i, src, ex, rs in sorted(lines, reverse=True)[:n]:
    top = sorted(rs.items(), key=lambda kv: -kv[1])[:2]
    print("%6d #%d ex=%d %s  %s" % (s, i, ex, src, top))
