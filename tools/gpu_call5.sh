#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn6.log 2>&1
rc=$?; echo "attn6 rc=$rc"; tail -12 gpurun_out/pytest_attn6.log
if [ $rc -ne 0 ]; then echo "v6 FAILED"; else
  for m in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py $m > gpurun_out/trace_attn6_$m.txt 2>&1; cat gpurun_out/trace_attn6_$m.txt; done
  for m in sd15 sdxl; do timeout 120 python tools/profile_kernels.py $m attn_self; done
  timeout 600 python -m pytest tests/test_processors_gpu.py tests/test_unet_gpu.py tests/test_clip_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet6.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet6.log
fi
