"""Summarise an `ncu --csv` per-launch duration list by kernel name: python tools/summarize_launches.py launches.csv [skip_first_n]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    rows.append((r[ki], ns))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for k, ns in rows:
    k = re.sub(r"\(.*", "", k)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (serialised, cold-cache: compare SHARES)")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns/1e6:9.3f} ms  {100*ns/tot:5.1f}%  x{n:<5d} {k}")
