"""scratch: summarise an `ncu --page source --csv` export: top SASS lines by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), int(r[iex]), i, r[isrc]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instr', sum(d[1] for d in data), 'lines', len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for d in sorted(data, reverse=True)[:n]:
    print("%6d %5.1f%% ex=%9d #%d %s" % (d[0], 100.0 * d[0] / tot, d[1], d[2], d[3][:120]))
