#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/splitk_sweep.txt
for sk in 0 2 3 4 6 8; do CID_GEMM_SPLITK=$sk CID_GEMM_SPLIT_MIN_KB=8 timeout 300 python tools/bench_splitk.py >> gpurun_out/splitk_sweep.txt 2>&1; done
cat gpurun_out/splitk_sweep.txt
