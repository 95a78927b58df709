#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_embed_gpu.py tests/test_checkpoint_gpu.py -x -q -s > gpurun_out/embed_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/embed_pytest.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or ln_rows or perceiver" > gpurun_out/embed_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/embed_kernels.log
for wl in sd15 sdxl; do
  timeout 600 ncu --kernel-name-base demangled -k regex:cid:: --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches_$wl.csv python tools/profile_step.py $wl 2 > gpurun_out/final_prof_$wl.log 2>&1
  timeout 600 ncu --kernel-name-base demangled -k regex:cid:: --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/final_dram_$wl.csv python tools/profile_step.py $wl 1 > gpurun_out/final_dram_$wl.log 2>&1
done
tail -25 gpurun_out/embed_pytest.log; tail -5 gpurun_out/embed_kernels.log; wc -l gpurun_out/final_launches_*.csv gpurun_out/final_dram_*.csv
