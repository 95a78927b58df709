#!/bin/bash
mkdir -p gpurun_out
for wl in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_cross.py $wl > gpurun_out/trace_cross2_$wl.txt 2>&1; cat gpurun_out/trace_cross2_$wl.txt | head -70; done
