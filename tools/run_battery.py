"""Run every kernel check of tests/kernel_checks.py in its own subprocess (with a timeout) and write
gpurun_out/battery.json.  Usage: python tools/run_battery.py [name-prefix ...]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.kernel_checks import CHECKS

def main():
    sel = sys.argv[1:]
    names = [n for n in CHECKS if not sel or any(n.startswith(s) for s in sel)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    env = dict(os.environ, PYTHONPATH=ROOT)
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "tests.kernel_checks", n], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                r = json.loads(line[-1][7:])
            else:
                r = {"name": n, "ok": False, "crash": True, "rc": p.returncode, "stdout": p.stdout[-1500:], "stderr": p.stderr[-2500:]}
        except subprocess.TimeoutExpired as e:
            r = {"name": n, "ok": False, "timeout": True, "stdout": (e.stdout or b"")[-1000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-1000:]}
        r["check"] = n
        r["secs"] = round(time.time() - t0, 2)
        results.append(r)
        print(("PASS " if r.get("ok") else "FAIL ") + n + "  " + json.dumps({k: r[k] for k in r if k in ("max_err", "tol", "ref_max", "nan", "crash", "timeout", "n_bad_rows", "n_bad_cols")}), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "battery.json"), "w") as f:
            json.dump(results, f, indent=1)
    nfail = sum(1 for r in results if not r.get("ok"))
    print(f"battery: {len(results) - nfail}/{len(results)} passed")
    return 1 if nfail else 0

if __name__ == "__main__":
    sys.exit(main())
