#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn5.log 2>&1
rc=$?; echo "attn5 rc=$rc"; tail -3 gpurun_out/pytest_attn5.log
if [ $rc -ne 0 ]; then export CID_LIB_PATH=$PWD/tools/bin/libcidb200_v4.so; echo "FALLBACK v4"; fi
if [ $rc -eq 0 ]; then
  for m in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py $m > gpurun_out/trace_attn5_$m.txt 2>&1; cat gpurun_out/trace_attn5_$m.txt; done
fi
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_kernels3.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/pytest_kernels3.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_processors_gpu.py tests/test_clip_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet3.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet3.log
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes3_$wl.txt 2>&1; grep -E "attn_self|layernorm|gn_apply|gn_stats" gpurun_out/shapes3_$wl.txt | head -12; done
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench3_$wl.json 2> gpurun_out/bench3_$wl.err; python -c "
import json;d=json.loads(open('gpurun_out/bench3_$wl.json').read().strip().splitlines()[-1]);print('$wl',d['value'],d['ms_per_step'],d['clocks'])"; done
