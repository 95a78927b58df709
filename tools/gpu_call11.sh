#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_checkpoint_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 120 python tools/profile_kernels.py sdxl gemm_out_proj_1280 gemm_plain_1280
for pol in "6,3" "4,5" "3,6"; do echo "splitk $pol"; CID_TOOL_SPLITK=$pol timeout 120 python tools/profile_kernels.py sdxl gemm_out_proj_1280 gemm_plain_1280 2>&1 | grep -v PROFILE; done
CID_TOOL_SPLITK="6,3" timeout 300 python tools/profile_shapes.py sdxl > gpurun_out/shapes11_sdxl_split63.txt 2>&1; head -8 gpurun_out/shapes11_sdxl_split63.txt
CID_TOOL_SPLITK="6,3" timeout 300 python tools/profile_shapes.py sd15 > gpurun_out/shapes11_sd15_split63.txt 2>&1; head -1 gpurun_out/shapes11_sd15_split63.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/ncu_outproj1280 python tools/profile_kernels.py sdxl gemm_out_proj_1280 > gpurun_out/ncu_outproj1280.log 2>&1
