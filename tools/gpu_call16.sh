#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -k attn_cross -q -p no:cacheprovider > gpurun_out/pytest_cross16.log 2>&1; rc=$?; echo "cross rc=$rc"; tail -5 gpurun_out/pytest_cross16.log
for wl in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_cross.py $wl > gpurun_out/trace_cross2c_$wl.txt 2>&1; head -36 gpurun_out/trace_cross2c_$wl.txt; done
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_processors_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet16.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet16.log
for wl in sd15 sdxl; do
  timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes16_$wl.txt 2>&1
  CID_LIB_PATH=$PWD/tools/bin/libcidb200_nostagger.so timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes16_nostagger_$wl.txt 2>&1
  grep -h attn_cross gpurun_out/shapes16_$wl.txt gpurun_out/shapes16_nostagger_$wl.txt
done
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench16_$wl.json 2> gpurun_out/bench16_$wl.err; python -c "
import json;d=json.loads(open('gpurun_out/bench16_$wl.json').read().strip().splitlines()[-1]);print('$wl',d['value'],d['ms_per_step'],d['clocks'])"; done
