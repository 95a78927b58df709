"""How does the time of ONE tile depend on how many SMs run tiles at the same time?  (a conv tile inside a full wave takes ~1.6x
as long as the same tile on an otherwise idle chip - which shared resource is it?)

Runs the 1280 -> 1280 3x3 convolution on 16x16 images for batch sizes that give 10 ... 148 concurrent 128x256 tiles (one per SM, tail
splitting off), with the L2 flushed between launches, and prints the launch time = time of one tile at that concurrency.  Follow up with
    ncu --set full -k regex:gemm_tc2 --launch-skip 3 -c 1 python tools/probe_wave_scaling.py 14      (full wave)
    ncu --set full -k regex:gemm_tc2 --launch-skip 3 -c 1 python tools/probe_wave_scaling.py 1       (10 tiles)
and compare lts / l1tex / fabric throughput and the tensor pipe.
    python tools/probe_wave_scaling.py [NB ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistentid_b200 import lib, ops

dev, dt = "cuda", torch.float16
lib.set_splitk(0, 1)                                     # whole tiles only
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
C, H = 1280, 16
w = torch.randn(C, 9 * C, device=dev, dtype=dt) * (9 * C) ** -0.5
b = torch.randn(C, device=dev, dtype=dt)
for NB in ([int(a) for a in sys.argv[1:]] or [1, 2, 3, 5, 7, 10, 12, 14]):
    x = torch.randn(NB, H, H, C, device=dev, dtype=dt)
    out = torch.empty(NB * H * H, C, device=dev, dtype=dt)
    fn = lambda: ops.conv3x3(x, w, out, NB, H, H, C, C, bias=b)
    for _ in range(3):
        fn()
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    tiles = (NB * H * H // 128) * (C // 256)
    t = ms[len(ms) // 2]
    print(f"NB={NB:3d} tiles={tiles:4d} (waves {tiles / 148:.2f})  launch {t * 1e3:8.1f} us   per-wave {t * 1e3 / max(1, -(-tiles // 148)):8.1f} us   "
          f"{2.0 * NB * H * H * C * 9 * C / t / 1e9:7.1f} TF/s", flush=True)
