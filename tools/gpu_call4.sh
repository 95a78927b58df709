#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn5b.log 2>&1; echo "attn5b rc=$?"; tail -2 gpurun_out/pytest_attn5b.log
for m in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py $m > gpurun_out/trace_attn5b_$m.txt 2>&1; cat gpurun_out/trace_attn5b_$m.txt; done
for m in sd15 sdxl; do timeout 120 python tools/profile_kernels.py $m attn_self; done
