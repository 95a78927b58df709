#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q > gpurun_out/r3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3_tests.log
for wl in sd15 sdxl; do CID_PDL=0 timeout 600 python tools/profile_shapes.py $wl > gpurun_out/shapes_$wl.txt 2>&1; done
for pdl in 0 1; do for wl in sd15 sdxl; do
  CID_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-profile > gpurun_out/pdlB${pdl}_$wl.json 2> gpurun_out/pdlB${pdl}_$wl.err
done; done
tail -3 gpurun_out/r3_tests.log
for f in gpurun_out/pdlB?_*.json; do echo $f; cut -c1-100 $f; done
head -30 gpurun_out/shapes_sd15.txt
