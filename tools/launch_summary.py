"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): last 1/k of the launches."""
import csv, sys
path, k = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
out = [(r[ik][:70], r[ig], float(r[iv].replace(",", "")) / 1000) for r in rows]
last = out[-(len(out) // k):]
for name, g, v in last:
    print(f"{v:9.1f} us  {g:18s} {name}")
print("sum us", round(sum(v for _, _, v in last), 1), "launches", len(last))
