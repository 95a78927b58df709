"""Print selected metrics of an .ncu-rep: python tools/ncu_metrics.py file.ncu-rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"gpu__time_duration.sum$|dram__bytes_(read|write).sum$|dram__bytes_(read|write).sum.per_second$|dram__throughput.*pct|lts__t_bytes.sum$|lts__throughput.avg.pct|sm__throughput.avg.pct|tensor.*(pct_of_peak_sustained_active|pct_of_peak_sustained_elapsed)$|warps_active.avg.pct|registers_per_thread$|issue_active.avg.pct|smsp__average_warp.*stall|smsp__warp_issue_stalled.*ratio|l1tex__data_pipe.*shared.*pct|smem|sm__inst_executed_pipe_(xu|fma|alu|fmaheavy|lsu|uniform).*pct_of_peak_sustained_active$|launch__(grid|block)_size|launch__occupancy_limit")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h) and v not in ("0", "0.000000", ""):
            print(f"  {h} = {v} {u}")
