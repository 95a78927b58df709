#!/bin/bash
# ncu --set full captures of the final kernels at BASELINE shapes (one launch each; never a bench number)
mkdir -p gpurun_out
cap() { # name model kernel-substr ncu-kernel-regex
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$4 --launch-skip 3 -c 1 -f -o gpurun_out/ncu_final_$1 python tools/profile_kernels.py $2 $3 > gpurun_out/ncu_final_$1.log 2>&1
}
cap conv_sd15 sd15 conv3x3 gemm_tc2
cap ff2_sdxl sdxl gemm_ff2 gemm_tc2
cap outproj_sdxl sdxl gemm_out_proj gemm_tc2
cap attn_self3_sd15 sd15 attn_self attn_self3
ls -la gpurun_out/ncu_final_*.ncu-rep
