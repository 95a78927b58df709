"""Time a few tail-heavy GEMM / conv shapes (CUDA events, L2-flushed between launches) for the current CID_GEMM_SPLITK / CID_GEMM_SPLIT_MIN_KB:
    CID_GEMM_SPLITK=4 python tools/bench_splitk.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import ops

dev, dt = "cuda", torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=12):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def gemm_case(M, N, K, residual=True):
    a = torch.randn(M, K, device=dev, dtype=dt); w = torch.randn(N, K, device=dev, dtype=dt) * K ** -0.5
    b = torch.randn(N, device=dev, dtype=dt); r = torch.randn(M, N, device=dev, dtype=dt) if residual else None
    out = torch.empty(M, N, device=dev, dtype=dt)
    return lambda: ops.gemm(a, w, out, bias=b, residual=r), 2.0 * M * N * K


def conv_case(NB, H, W, Cin, Cout):
    x = torch.randn(NB, H, W, Cin, device=dev, dtype=dt); w = torch.randn(Cout, 9 * Cin, device=dev, dtype=dt) * (9 * Cin) ** -0.5
    b = torch.randn(Cout, device=dev, dtype=dt); out = torch.empty(NB * H * W, Cout, device=dev, dtype=dt)
    return lambda: ops.conv3x3(x, w, out, NB, H, W, Cin, Cout, bias=b), 2.0 * NB * H * W * Cout * 9 * Cin


cases = {"gemm 4096x1280x5120": gemm_case(4096, 1280, 5120), "gemm 8192x1280x5120": gemm_case(8192, 1280, 5120),
         "gemm 4096x1280x2560": gemm_case(4096, 1280, 2560), "gemm 1024x1280x5120": gemm_case(1024, 1280, 5120),
         "conv 16x16x16 1280->1280": conv_case(16, 16, 16, 1280, 1280), "conv 16x8x8 1280->1280": conv_case(16, 8, 8, 1280, 1280),
         "conv 8x32x32 1280->1280": conv_case(8, 32, 32, 1280, 1280), "conv 16x16x16 2560->1280": conv_case(16, 16, 16, 2560, 1280)}
tag = f"SPLITK={os.environ.get('CID_GEMM_SPLITK', 'dflt')} MIN_KB={os.environ.get('CID_GEMM_SPLIT_MIN_KB', 'dflt')}"
for name, (fn, fl) in cases.items():
    ms = timeit(fn)
    print(f"{tag:24s} {name:28s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TF/s", flush=True)
