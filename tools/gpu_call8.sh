#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn7.log 2>&1; rc=$?; echo "attn7 rc=$rc"; tail -8 gpurun_out/pytest_attn7.log
if [ $rc -ne 0 ]; then exit 1; fi
CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py sd15 > gpurun_out/trace_attn7_sd15.txt 2>&1; head -12 gpurun_out/trace_attn7_sd15.txt
for m in sd15 sdxl; do timeout 120 python tools/profile_kernels.py $m attn_self; CID_LIB_PATH=$PWD/tools/bin/libcidb200_v6.so timeout 120 python tools/profile_kernels.py $m attn_self; done
timeout 600 python -m pytest tests/test_processors_gpu.py tests/test_unet_gpu.py tests/test_clip_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet8.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet8.log
