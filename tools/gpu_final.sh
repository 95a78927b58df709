#!/bin/bash
# Final validation + evidence run (one GPU): whole -m gpu suite, smoke, driver-style bench, ncu launch lists and full captures.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_final.log; tail -4 gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 1500 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; echo "ref rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for wl in sd15 sdxl; do
  timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_step_$wl.csv python tools/profile_step.py $wl 1 > gpurun_out/ncu_step_$wl.log 2>&1
  python tools/summarize_ncu_step.py gpurun_out/ncu_step_$wl.csv > gpurun_out/ncu_step_${wl}_summary.txt 2>&1; head -14 gpurun_out/ncu_step_${wl}_summary.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_self7 -s 3 -c 1 -o gpurun_out/ncu_attn7_sd15 python tools/profile_kernels.py sd15 attn_self > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_self7 -s 3 -c 1 -o gpurun_out/ncu_attn7_sdxl python tools/profile_kernels.py sdxl attn_self > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/ncu_conv_sd15 python tools/profile_kernels.py sd15 conv3x3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/ncu_ff1_sdxl python tools/profile_kernels.py sdxl gemm_ff1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
tail -c 600 gpurun_out/bench_final.json
