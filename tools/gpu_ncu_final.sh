#!/bin/bash
# ncu --set full captures of the final kernels at BASELINE shapes (one launch each; never a bench number).
# The reports are reduced to text ON THE BOX (metric extract + top stall lines); the .ncu-rep files are too big to travel back.
mkdir -p gpurun_out /tmp/ncu
cap() { # name model kernel-substr ncu-kernel-regex
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$4 --launch-skip 3 -c 1 -f -o /tmp/ncu/$1 python tools/profile_kernels.py $2 $3 > /tmp/ncu/$1.log 2>&1
  python tools/ncu_metrics.py /tmp/ncu/$1.ncu-rep > gpurun_out/ncu_final_$1_metrics.txt 2>&1
  python tools/ncu_hot.py /tmp/ncu/$1.ncu-rep 30 > gpurun_out/ncu_final_$1_hot.txt 2>&1
}
cap conv_sd15 sd15 conv3x3 gemm_tc2
cap ff2_sdxl sdxl gemm_ff2 gemm_tc2
cap outproj_sdxl sdxl gemm_out_proj gemm_tc2
cap attn_self3_sd15 sd15 attn_self attn_self3
wc -l gpurun_out/ncu_final_*
