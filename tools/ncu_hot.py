"""Top stall-sample SASS lines of an .ncu-rep source page: python tools/ncu_hot.py file.ncu-rep [topn]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
si, ai, ns = hdr.index("Source"), hdr.index("Address"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[1:]:
    try:
        n = int(r[ns])
    except Exception:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print("total samples", tot)
for idx, (n, r) in enumerate(sorted(data, key=lambda x: -x[0])[:topn]):
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:3]
    print(f"{100*n/tot:5.1f}%  {r[si][:95]:95s} {[(h[6:], c) for c, h in st if c]}")
