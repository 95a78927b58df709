"""SASS evidence of the Blackwell-native paths: per kernel of libcidb200.so, how many tcgen05 / TMEM / TMA instructions the
compiled code holds.   python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt
Mnemonics (guide B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG, tcgen05.commit -> UTCBAR,
mma.sync would show as HMMA (none expected)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "consistentid_b200", "libcidb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = {"UTCHMMA": r"\bUTCHMMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG", "UTCBAR": r"\bUTCBAR",
       "SYNCS": r"\bSYNCS", "MUFU.EX2": r"\bMUFU\.EX2", "HMMA": r"\bHMMA", "SHFL": r"\bSHFL", "RED/ATOM": r"\b(RED|ATOMG|ATOMS)\b"}
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("cid::", "")
        rows[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    rows[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", line) else 0
    for k, p in PAT.items():
        if re.search(p, line):
            rows[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass); arch sm_100a")
cols = ["instructions"] + list(PAT)
print(f"{'kernel':58s} " + " ".join(f"{c:>9s}" for c in cols))
tot = collections.Counter()
for k, c in rows.items():
    print(f"{k[:58]:58s} " + " ".join(f"{c[x]:9d}" for x in cols))
    tot.update(c)
print(f"{'TOTAL':58s} " + " ".join(f"{tot[x]:9d}" for x in cols))
