#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -k "attn_cross" -q -p no:cacheprovider > gpurun_out/pytest_k18.log 2>&1; rc=$?; echo "kernels rc=$rc"; tail -4 gpurun_out/pytest_k18.log
CID_LIB_PATH=$PWD/tools/bin/libcidb200_gtrace.so timeout 300 python tools/trace_gemm.py > gpurun_out/trace_gemm18.txt 2>&1; cat gpurun_out/trace_gemm18.txt
for wl in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_cross.py $wl > gpurun_out/trace_cross2e_$wl.txt 2>&1; head -40 gpurun_out/trace_cross2e_$wl.txt; done
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes18_$wl.txt 2>&1; grep -h attn_cross gpurun_out/shapes18_$wl.txt; done
