"""Launch the hot kernels at BASELINE shapes a few times (for `ncu` captures and quick CUDA-event timing).
usage: python tools/profile_kernels.py [sd15|sdxl] [kernel-name-substr ...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import ops, lib
from consistentid_b200.weights import interleave_geglu

def rnd(shape, dt, s=1.0):
    return (torch.randn(shape, device="cuda") * s).to(dt)

def cases(model):
    dt = torch.float16 if model == "sd15" else torch.bfloat16
    out = {}
    if model == "sd15":
        NB, H, W, C, heads = 16, 64, 64, 320, 8
    else:
        NB, H, W, C, heads = 8, 64, 64, 640, 10
    M, d = NB * H * W, C // heads
    x = rnd((M, C), dt); res = rnd((M, C), dt); o = torch.empty((M, C), dtype=dt, device="cuda")
    w = rnd((C, C), dt, C ** -0.5); b = rnd((C,), dt)
    out["gemm_out_proj"] = (lambda: ops.gemm(x, w, o, bias=b, residual=res), 2.0 * M * C * C)
    if model == "sdxl":     # the level-2 workhorse: 8192 x 1280 x 1280, residual stream updated in place, LayerNorm row statistics in the epilogue
        M2, C2 = 8192, 1280
        x2 = rnd((M2, C2), dt); t2 = rnd((M2, C2), dt); w2_ = rnd((C2, C2), dt, C2 ** -0.5); b2 = rnd((C2,), dt)
        rs2 = torch.zeros((M2, 2), dtype=torch.float32, device="cuda")
        out["gemm_out_proj_1280"] = (lambda: ops.gemm(x2, w2_, t2, bias=b2, residual=t2, row_stats=rs2), 2.0 * M2 * C2 * C2)
        out["gemm_plain_1280"] = (lambda: ops.gemm(x2, w2_, t2, bias=b2), 2.0 * M2 * C2 * C2)
    wqkv = rnd((3 * C, C), dt, C ** -0.5); qk = torch.empty((M, 2 * C), dtype=dt, device="cuda"); vt = torch.empty((NB * heads, d, H * W), dtype=dt, device="cuda")
    out["gemm_qkv"] = (lambda: ops.gemm(x, wqkv, qk, epi=lib.EPI_QKV, vt=vt, n_split=2 * C, heads=heads, hdim=d, ntok=H * W), 2.0 * M * 3 * C * C)
    w1 = rnd((8 * C, C), dt, C ** -0.5); b1 = rnd((8 * C,), dt)
    w1i, b1i = interleave_geglu(w1, b1, lib.gemm_tile_n(8 * C, lib.EPI_GEGLU)); ffm = torch.empty((M, 4 * C), dtype=dt, device="cuda")
    out["gemm_ff1_geglu"] = (lambda: ops.gemm(x, w1i, ffm, bias=b1i, epi=lib.EPI_GEGLU), 2.0 * M * 8 * C * C)
    w2 = rnd((C, 4 * C), dt, (4 * C) ** -0.5)
    out["gemm_ff2"] = (lambda: ops.gemm(ffm, w2, o, bias=b, residual=res), 2.0 * M * 4 * C * C)
    wc = rnd((C, 9 * C), dt, (9 * C) ** -0.5); xi = rnd((NB, H, W, C), dt)
    out["conv3x3"] = (lambda: ops.conv3x3(xi, wc, o, NB, H, W, C, C, bias=b, residual=res), 2.0 * M * C * 9 * C)
    qk.normal_(); vt.normal_()
    out["attn_self"] = (lambda: ops.attn_self(qk[:, :C], qk[:, C:], vt, o, NB, heads, H * W, d), 4.0 * NB * heads * (H * W) ** 2 * d)
    kc = rnd((NB, 96, C), dt); vc = rnd((NB * heads, d, 96), dt); q = rnd((M, C), dt)
    out["attn_cross"] = (lambda: ops.attn_cross(q, kc, vc, o, NB, heads, H * W, d, 77, 4, 1.0), 4.0 * NB * heads * H * W * 81 * d)
    return out

if __name__ == "__main__":
    if os.environ.get("CID_TOOL_SPLITK"):          # tail-balancing policy override for experiments: "max_split,min_kblocks"
        lib.set_splitk(*[int(v) for v in os.environ["CID_TOOL_SPLITK"].split(",")])
    model = sys.argv[1] if len(sys.argv) > 1 else "sd15"
    sel = sys.argv[2:]
    res = {}
    for name, (fn, flops) in cases(model).items():
        if sel and not any(s in name for s in sel):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))
        print(name, res[name], flush=True)
    print("PROFILE_KERNELS " + json.dumps({model: res}))
