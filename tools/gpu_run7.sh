#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r7_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r7_kernels.log
for m in sd15 sdxl; do timeout 300 python tools/profile_kernels.py $m gemm conv > gpurun_out/r7_kern_$m.log 2>&1; done
timeout 300 python tools/bench_splitk.py > gpurun_out/r7_shapes.txt 2>&1
for wl in sd15 sdxl; do timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-profile > gpurun_out/r7_$wl.json 2> gpurun_out/r7_$wl.err; done
tail -2 gpurun_out/r7_kernels.log; grep -h "gemm_\|conv3x3" gpurun_out/r7_kern_sd15.log gpurun_out/r7_kern_sdxl.log | tail -12; cat gpurun_out/r7_shapes.txt | cut -c25-; for wl in sd15 sdxl; do cut -c1-100 gpurun_out/r7_$wl.json; done
