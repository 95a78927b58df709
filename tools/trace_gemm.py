"""Per-tile phase timeline of gemm_tc2_kernel (debug build):
    tools/build_variant.sh gtrace -DCID_GEMM_TRACE && CID_LIB_PATH=tools/bin/libcidb200_gtrace.so python tools/trace_gemm.py [case ...]
cases: out320 (65536x320x320 +bias +residual, fp16), out320_ln (+ row_stats producer), q320 (folded-LayerNorm consumer), out640 (16384x640x640),
out1280 (8192x1280x1280 +residual +row_stats, bf16), plain1280, ff2_1280 (8192x1280x5120).
Epilogue thread 0 stamps per tile: 0 start | 1 staging loops done | 2 prev TMA store drained + sync | 3 residual staged + sync | 4 accumulator ready |
5 drained | 6 final sync | 7 released.  MMA warp: 0 start | 1 accumulator free | 2 first k-block landed | 3 last MMA issued.
Producer: 0 start | 1 first slot free | 2 last TMA issued."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import lib, ops

def rnd(shape, dt, s=1.0):
    return (torch.randn(shape, device="cuda") * s).to(dt)

def make(case):
    f16, b16 = torch.float16, torch.bfloat16
    M, N, K, dt, kw = {"out320": (65536, 320, 320, f16, "res"), "out320_ln": (65536, 320, 320, f16, "res+rs"), "q320": (65536, 320, 320, f16, "ln"),
                       "plain320": (65536, 320, 320, f16, ""), "out640": (16384, 640, 640, f16, "res+rs"), "out1280": (8192, 1280, 1280, b16, "res+rs"),
                       "plain1280": (8192, 1280, 1280, b16, ""), "ff2_1280": (8192, 1280, 5120, b16, "res+rs")}[case]
    x = rnd((M, K), dt); w = rnd((N, K), dt, K ** -0.5); b = rnd((N,), dt); o = rnd((M, N), dt)
    args = dict(bias=b)
    if "res" in kw: args["residual"] = o
    if "rs" in kw: args["row_stats"] = torch.zeros((M, 2), dtype=torch.float32, device="cuda")
    if "ln" in kw:
        st = torch.stack([x.float().sum(1), (x.float() ** 2).sum(1)], 1).contiguous()
        args["ln"] = (st, w.float().sum(1).contiguous(), 1e-5)
    return (lambda: ops.gemm(x, w, o, **args)), 2.0 * M * N * K, (M, N, K)

trace = torch.zeros(3 * 16 * 64 * 8, dtype=torch.int64, device="cuda")
lib._lib.cid_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
for case in (sys.argv[1:] or ["out320", "q320", "plain320", "out640", "out1280", "plain1280", "ff2_1280"]):
    fn, flops, shape = make(case)
    lib._lib.cid_debug_set_gemm_trace(None)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    trace.zero_(); lib._lib.cid_debug_set_gemm_trace(trace.data_ptr()); fn(); torch.cuda.synchronize(); lib._lib.cid_debug_set_gemm_trace(None)
    t = trace.cpu().view(3, 16, 64, 8)
    print(f"== {case} {shape}: {us:.1f} us/launch, {flops / us / 1e6:.0f} TFLOP/s")
    for cta in (0, 5):
        ep, mm, pr = t[0, cta], t[1, cta], t[2, cta]
        n = int((ep[:, 0] != 0).sum()); base = int(pr[0, 0]) if int(pr[0, 0]) else int(ep[0, 0])
        print(f"  cta {cta}: {n} tiles.  epilogue: start | staging | TMA drain+sync | res staged+sync | wait acc | drain | final sync | release || MMA: start | wait acc free | wait 1st kb | issue all || producer: start | 1st slot wait | issue all")
        for it in range(min(n, 8)):
            e, m, p = ep[it], mm[it], pr[it]
            print(f"   tile {it}: {int(e[0]) - base:7d} | " + " | ".join(f"{int(e[k + 1] - e[k]):5d}" for k in range(7))
                  + f" || {int(m[0]) - base:7d} | " + " | ".join(f"{int(m[k + 1] - m[k]):5d}" for k in range(3))
                  + f" || {int(p[0]) - base:7d} | " + " | ".join(f"{int(p[k + 1] - p[k]):5d}" for k in range(2)))
