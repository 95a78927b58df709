#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn7b.log 2>&1; rc=$?; echo "attn7b rc=$rc"; tail -3 gpurun_out/pytest_attn7b.log
CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py sd15 > gpurun_out/trace_attn7b_sd15.txt 2>&1; head -14 gpurun_out/trace_attn7b_sd15.txt
for m in sd15 sdxl; do timeout 120 python tools/profile_kernels.py $m attn_self; done
