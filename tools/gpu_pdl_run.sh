#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_embed_gpu.py tests/test_checkpoint_gpu.py -x -q -s > gpurun_out/embed_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/embed_pytest.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/pdl_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/pdl_kernels.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py -x -q > gpurun_out/pdl_unet.log 2>&1; echo "rc=$?" >> gpurun_out/pdl_unet.log
for pdl in 0 1; do for wl in sd15 sdxl; do
  CID_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-profile > gpurun_out/pdl${pdl}_$wl.json 2> gpurun_out/pdl${pdl}_$wl.err
done; done
tail -4 gpurun_out/embed_pytest.log; tail -3 gpurun_out/pdl_kernels.log; tail -3 gpurun_out/pdl_unet.log
for f in gpurun_out/pdl?_*.json; do echo $f; cut -c1-120 $f; done
