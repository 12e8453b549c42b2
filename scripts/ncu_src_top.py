"""Summarise an `ncu --page source --csv --print-source sass` export: hottest SASS lines by stall samples.
usage: ncu_src_top.py file.csv [min_samples] """
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("total samples", tot, "lines", len(data))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        agg[s] += int(r[col[s]] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, r in enumerate(data):
    n = int(r[col["# Samples"]] or 0)
    if n < mn:
        continue
    ex = int(r[col["Instructions Executed"]] or 0)
    top = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print("%5d %7d %10d  %-70s %s" % (i, n, ex, r[col["Source"]].strip()[:70], " ".join("%s=%d" % (s, v) for v, s in top if v)))
