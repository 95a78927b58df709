"""Per-tile phase timeline of the self-attention softmax warps (debug build):
    tools/build_variant.sh trace -DCID_ATTN_TRACE && CID_LIB_PATH=tools/bin/libcidb200_trace.so python tools/trace_attn.py [sd15|sdxl]
Prints, for a few CTAs, the average cycles per key tile spent in each phase of softmax warp 2 / lane 0:
  wait S | tcgen05.ld S | row max (+ rare rescale) | wait P buffer | exp + store P | fence + arrive."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import lib, ops
model = sys.argv[1] if len(sys.argv) > 1 else "sd15"
dt = torch.float16 if model == "sd15" else torch.bfloat16
NB, H, N, d = (16, 8, 4096, 40) if model == "sd15" else (8, 10, 4096, 64)
C = H * d
qk = torch.randn(NB * N, 2 * C, device="cuda").to(dt); vt = torch.randn(NB * H, d, N, device="cuda").to(dt)
o = torch.empty(NB * N, C, device="cuda", dtype=dt)
trace = torch.zeros(64 * 64 * 8, dtype=torch.int64, device="cuda")
lib._lib.cid_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attn_self(qk[:, :C], qk[:, C:], vt, o, NB, H, N, d)
lib._lib.cid_debug_set_attn_trace(trace.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attn_self(qk[:, :C], qk[:, C:], vt, o, NB, H, N, d); e1.record(); torch.cuda.synchronize()
print(f"{model}: kernel {e0.elapsed_time(e1) * 1e3:.0f} us")
t = trace.cpu().view(64, 64, 8)
T = N // 128
# v5: rows = CTAs (one 128-query tile each); v6: rows = (CTA, warpgroup) pairs (two tiles per CTA) and phase 3->4 is the wait for the MUFU token
names = ["wait S", "ld S", "max", "exp (v7) | wait Pbuf/token (v5/v6)", "wait P.V (v7) | exp", "st P + arrive"]
print("row smid | per-tile period | " + " | ".join(names))
for cta in list(range(0, 8)) + list(range(8, min(32, N // 128), 4)):
    tt = t[cta, :T]
    smid = int(t[cta, 0, 7])
    ph = [(tt[2:, k + 1] - tt[2:, k]).float().mean().item() for k in range(6)]
    period = (tt[3:, 0] - tt[2:-1, 0]).float().mean().item()
    print(f"{cta:3d} {smid:4d} | {period:8.0f} | " + " | ".join(f"{p:7.0f}" for p in ph))
print("tile-by-tile (cta 0), cycles since tile 2 start: " + " ".join(str(int(t[0, j, 0] - t[0, 2, 0])) for j in range(2, 12)))

# v6 only: the MMA thread of CTA 0 (trace rows 32 + blockIdx.x): when it saw P_i, finished issuing P.V_i, finished issuing S_i(j+1),
# next to the softmax warpgroups' own stamps (same SM clock)
if t[32, 3].abs().sum() > 0:
    m, w0, w1 = t[32], t[0], t[1]
    base = int(w0[3, 0])
    print("CTA 0 timeline (cycles since warpgroup 0 entered tile 3); P = p_full seen, PV = P.V issued, S = S_i(j+1) issued, Sdone = softmax saw s_full")
    for j in range(3, 7):
        rel = lambda x: int(x) - base
        print(f"  tile {j}: WG0 exp [{rel(w0[j, 4])}, {rel(w0[j, 5])}] arrive {rel(w0[j, 6])} | MMA P0 {rel(m[j, 0])} PV0 {rel(m[j, 1])} S0 {rel(m[j, 2])} | WG0 next Sdone {rel(w0[j + 1, 1])}"
              f" || WG1 exp [{rel(w1[j, 4])}, {rel(w1[j, 5])}] arrive {rel(w1[j, 6])} | MMA P1 {rel(m[j, 3])} PV1 {rel(m[j, 4])} S1 {rel(m[j, 5])} | WG1 next Sdone {rel(w1[j + 1, 1])}")
