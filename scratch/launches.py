"""scratch: print (kernel, grid, us) from an ncu --csv launch list"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    if 'at::' in d["Kernel Name"]: continue
    print("%-62s %-16s %10.1f us" % (d["Kernel Name"][:62], d["Grid Size"], float(d["Metric Value"]) / 1e3))
