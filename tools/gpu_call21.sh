#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_k21.log 2>&1; rc=$?; echo "kernels rc=$rc"; tail -6 gpurun_out/pytest_k21.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_checkpoint_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet21.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet21.log
python - <<'P'
import torch, sys
sys.path.insert(0, '.')
from consistentid_b200 import ops
dt = torch.float16
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for NB, HW, C in ((16, 64, 1280), (16, 64, 2560), (16, 256, 1280), (16, 256, 2560), (16, 1024, 640)):
    x = torch.randn(NB, HW, C, device='cuda').to(dt); g = torch.randn(C, device='cuda').to(dt); b = torch.randn(C, device='cuda').to(dt)
    out = torch.empty_like(x); sums = torch.zeros(NB, 32, 2, device='cuda'); ch = torch.randn(NB, C, 2, device='cuda').abs()
    r = {}
    if ops.gn_small_ok(C, 0, HW, 32): r['small'] = t(lambda: ops.gn_small(x, C, None, 0, NB, HW, 32, g, b, 1e-5, True, out))
    r['stats+apply'] = t(lambda: (ops.gn_stats(x, C, None, 0, NB, HW, 32, sums), ops.gn_apply(x, C, None, 0, NB, HW, 32, sums, g, b, 1e-5, True, out)))
    r['apply_ch'] = t(lambda: ops.gn_apply_ch(x, C, ch, None, 0, None, NB, HW, 32, g, b, 1e-5, True, out))
    print(NB, HW, C, {k: round(v, 1) for k, v in r.items()}, 'us')
P
timeout 300 python tools/profile_shapes.py sd15 > gpurun_out/shapes21_sd15.txt 2>&1; head -1 gpurun_out/shapes21_sd15.txt; grep -E "^gn_" gpurun_out/shapes21_sd15.txt
timeout 400 python bench.py --workload sd15 --steps 4 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench21_sd15.json 2> gpurun_out/bench21_sd15.err; python -c "
import json;d=json.loads(open('gpurun_out/bench21_sd15.json').read().strip().splitlines()[-1]);print('sd15',d['value'],d['ms_per_step'],d['launches_per_denoise_step'],d['clocks'])"
