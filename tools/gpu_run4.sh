#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r4_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r4_kernels.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_processors_gpu.py -x -q > gpurun_out/r4_unet.log 2>&1; echo "rc=$?" >> gpurun_out/r4_unet.log
for sk in 0 8; do for wl in sd15 sdxl; do
  CID_GEMM_SPLITK=$sk timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-profile > gpurun_out/sk${sk}_$wl.json 2> gpurun_out/sk${sk}_$wl.err
done; done
for wl in sd15 sdxl; do timeout 600 python tools/profile_shapes.py $wl > gpurun_out/shapes_sk_$wl.txt 2>&1; done
tail -3 gpurun_out/r4_kernels.log; tail -3 gpurun_out/r4_unet.log
for f in gpurun_out/sk?_*.json; do echo $f; cut -c1-100 $f; done
head -14 gpurun_out/shapes_sk_sd15.txt
