"""Per-kernel summary of an `ncu --csv` launch list that carries, per launch, gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
dram__throughput.avg.pct_of_peak_sustained_elapsed and sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed:
    python tools/summarize_ncu_step.py step.csv [hbm_peak_GBs]
-> launches, time and share of the step, DRAM bytes, ACHIEVED GB/s (dram bytes / duration) and its fraction of the measured copy bandwidth,
   ncu's own dram-throughput % and tensor-pipe % (time-weighted).  ncu times are cold-cache and serialised: compare shares, not absolutes."""
import csv, sys, re, json, collections, os
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "%": 1.0}
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6572.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if len(sys.argv) <= 2 else peak
except Exception:
    pass
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.OrderedDict()
for r in rd:
    try:
        v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1)
    except ValueError:
        continue
    e = per.setdefault(r[ii], dict(k=re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cid::", ""), us=0.0, rd=0.0, wr=0.0, dpct=0.0, tpct=0.0))
    m = r[mi]
    if m.startswith("gpu__time_duration"): e["us"] = v
    elif m.startswith("dram__bytes_read"): e["rd"] = v
    elif m.startswith("dram__bytes_write"): e["wr"] = v
    elif m.startswith("dram__throughput"): e["dpct"] = v
    elif m.startswith("sm__pipe_tensor_cycles_active"): e["tpct"] = v
agg = collections.OrderedDict()
for e in per.values():
    a = agg.setdefault(e["k"], dict(n=0, us=0.0, by=0.0, dw=0.0, tw=0.0))
    a["n"] += 1; a["us"] += e["us"]; a["by"] += e["rd"] + e["wr"]; a["dw"] += e["dpct"] * e["us"]; a["tw"] += e["tpct"] * e["us"]
tot = sum(a["us"] for a in agg.values())
print(f"# {sys.argv[1]}: {len(per)} launches, {tot/1e3:.3f} ms serialised; achieved GB/s = ncu DRAM bytes / ncu duration; HBM peak {peak:.0f} GB/s (measured copy)")
print(f"{'kernel':44s} {'n':>5s} {'ms':>9s} {'share':>6s} {'DRAM MB':>10s} {'GB/s':>7s} {'of HBM':>7s} {'ncu dram%':>9s} {'tensor%':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = a["by"] / max(a["us"], 1e-9) / 1e3
    print(f"{k[:44]:44s} {a['n']:5d} {a['us']/1e3:9.3f} {100*a['us']/tot:5.1f}% {a['by']/1e6:10.1f} {gbs:7.0f} {gbs/peak:7.2f} {a['dw']/max(a['us'],1e-9):9.1f} {a['tw']/max(a['us'],1e-9):8.1f}")
