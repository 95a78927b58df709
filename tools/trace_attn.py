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
names = ["wait S", "ld S", "max", "wait Pbuf/token", "exp", "st P + arrive"]
print("row smid | per-tile period | " + " | ".join(names))
for cta in list(range(0, 8)) + list(range(8, min(32, N // 128), 4)):
    tt = t[cta, :T]
    smid = int(t[cta, 0, 7])
    ph = [(tt[2:, k + 1] - tt[2:, k]).float().mean().item() for k in range(6)]
    period = (tt[3:, 0] - tt[2:-1, 0]).float().mean().item()
    print(f"{cta:3d} {smid:4d} | {period:8.0f} | " + " | ".join(f"{p:7.0f}" for p in ph))
print("tile-by-tile (cta 0), cycles since tile 2 start: " + " ".join(str(int(t[0, j, 0] - t[0, 2, 0])) for j in range(2, 12)))
