#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_kernels10.log 2>&1; rc=$?; echo "kernels rc=$rc"; tail -6 gpurun_out/pytest_kernels10.log
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_processors_gpu.py tests/test_controlnet_gpu.py tests/test_fullsize_gpu.py tests/test_checkpoint_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet10.log 2>&1; echo "unet rc=$?"; tail -6 gpurun_out/pytest_unet10.log
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes10_$wl.txt 2>&1; head -1 gpurun_out/shapes10_$wl.txt; done
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench10_$wl.json 2> gpurun_out/bench10_$wl.err; python -c "
import json;d=json.loads(open('gpurun_out/bench10_$wl.json').read().strip().splitlines()[-1]);print('$wl',d['value'],d['ms_per_step'],d['clocks'], d['launches_per_denoise_step'])"; done
