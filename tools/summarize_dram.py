"""Summarise an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum` launch list by kernel name and write the
per-launch DRAM traffic of the tensor-core kernels (GEMM + implicit-GEMM conv) as JSON:
    python tools/summarize_dram.py dram.csv out.json
bench.py reads profiles/r01_dram_traffic_<workload>.json to fill roofline.traffic."""
import csv, sys, re, json, collections

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.OrderedDict()          # launch id -> [kernel, read, write]
for r in rd:
    v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1)
    e = per.setdefault(r[ii], [re.sub(r"\(.*", "", r[ki]), 0.0, 0.0])
    if "read" in r[mi]:
        e[1] += v
    elif "write" in r[mi]:
        e[2] += v
agg = collections.OrderedDict()
for k, rdb, wrb in per.values():
    a = agg.setdefault(k, dict(launches=0, read=0.0, write=0.0))
    a["launches"] += 1; a["read"] += rdb; a["write"] += wrb
tc = dict(launches=0, read=0.0, write=0.0)
for k, a in agg.items():
    if "gemm_tc" in k:
        for f in tc:
            tc[f] += a[f]
out = {"source": sys.argv[1], "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one eager denoising iteration",
       "tensor_core_kernels": dict(launches=tc["launches"], bytes_per_launch=(tc["read"] + tc["write"]) / max(tc["launches"], 1),
                                   read_bytes=tc["read"], write_bytes=tc["write"]),
       "by_kernel": {k: dict(launches=a["launches"], read_MB=round(a["read"] / 1e6, 2), write_MB=round(a["write"] / 1e6, 2)) for k, a in agg.items()}}
json.dump(out, open(sys.argv[2], "w"), indent=1)
tot = sum(a["read"] + a["write"] for a in agg.values())
print(f"launches {len(per)}  dram total {tot/1e6:.1f} MB")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"])):
    print(f"{(a['read']+a['write'])/1e6:10.1f} MB  (rd {a['read']/1e6:9.1f} wr {a['write']/1e6:9.1f})  x{a['launches']:<5d} {k}")
