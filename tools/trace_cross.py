"""Per-unit phase timeline of attn_cross2_kernel (debug build):
    tools/build_variant.sh trace -DCID_ATTN_TRACE && CID_LIB_PATH=tools/bin/libcidb200_trace.so python tools/trace_cross.py [sd15|sdxl]
Softmax warpgroup stamps (warp quarter 0, lane 0) per unit: 0 loop top, 1 S ready, 2 S in registers, 3 row maxima, 4 exponentials + pack,
5 P stored + arrive, 6 O ready, 7 O drained + stored.  MMA warp stamps per unit x: 0 loop top, 1 P.V(x) issued, 2 S(x+3) issued."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import lib, ops
model = sys.argv[1] if len(sys.argv) > 1 else "sd15"
dt = torch.float16 if model == "sd15" else torch.bfloat16
B, H, N, d = (16, 8, 4096, 40) if model == "sd15" else (8, 20, 1024, 64)
C = H * d
q = torch.randn(B * N, C, device="cuda").to(dt)
kt, vt_ = torch.randn(B * 77, C, device="cuda").to(dt), torch.randn(B * 77, C, device="cuda").to(dt)
ki, vi = torch.randn(B * 4, C, device="cuda").to(dt), torch.randn(B * 4, C, device="cuda").to(dt)
k_cat = torch.zeros(B, 96, C, dtype=dt, device="cuda"); vt_cat = torch.zeros(B * H, d, 96, dtype=dt, device="cuda")
ops.pack_cross_kv(kt, vt_, ki, vi, k_cat, vt_cat, B, C, H, 77, 4)
o = torch.empty(B * N, C, device="cuda", dtype=dt)
trace = torch.zeros(64 * 64 * 8, dtype=torch.int64, device="cuda")
lib._lib.cid_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
run = lambda: ops.attn_cross(q, k_cat, vt_cat, o, B, H, N, d, 77, 4, 1.0)
for _ in range(2): run()
lib._lib.cid_debug_set_attn_trace(trace.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"{model}: kernel {e0.elapsed_time(e1) * 1e3:.0f} us")
t = trace.cpu().view(64, 64, 8)
for cta in (0, 1, 7):
    for wg in (0, 1, 2):
        row = t[cta * 3 + wg]
        n = int((row[:, 0] != 0).sum())
        base = int(t[cta * 3, 0, 0])
        print(f"cta {cta} wg {wg}: {n} units; per unit: start | wait S | ld S | max | exp | st P | wait O | drain")
        for k in range(min(n, 8)):
            r = row[k]
            print(f"   unit {k}: {int(r[0]) - base:7d} | " + " | ".join(f"{int(r[e + 1] - r[e]):6d}" for e in range(7)))
    m = t[32 + cta]; base = int(t[cta * 3, 0, 0])
    n = int((m[:, 0] != 0).sum())
    print(f"cta {cta} MMA warp: unit x: loop top (rel. to wg0 start) | wait P(x) + issue P.V(x) | wait Q/KV + issue S(x+3) | -")
    for u in range(min(n, 12)):
        r = m[u]
        print(f"   unit {u}: {int(r[0]) - base:7d} | " + " | ".join(f"{int(r[e + 1] - r[e]):6d}" for e in range(3)))
