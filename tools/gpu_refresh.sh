#!/bin/bash
# Re-run of the validation + bench part of tools/gpu_final.sh on the final build (after the one-pass small GroupNorm landed)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 1500 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
KREG='regex:gemm_tc2|attn_|gn_|layernorm|upsample|phase_split|cfg_sched|skinny|pack_cross|latents_to|advance_step|timestep_embed|nchw|rows_to'
for wl in sd15; do
  timeout 600 ncu --metrics $M --clock-control none -k "$KREG" --csv --log-file gpurun_out/ncu_step_$wl.csv python tools/profile_step.py $wl 1 > gpurun_out/ncu_step_$wl.log 2>&1
  python tools/summarize_ncu_step.py gpurun_out/ncu_step_$wl.csv > gpurun_out/ncu_step_${wl}_summary.txt 2>&1; head -12 gpurun_out/ncu_step_${wl}_summary.txt
  python tools/summarize_dram.py gpurun_out/ncu_step_$wl.csv gpurun_out/dram_traffic_$wl.json > gpurun_out/dram_${wl}_summary.txt 2>&1
  gzip -f gpurun_out/ncu_step_$wl.csv
done
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes_final_$wl.txt 2>&1; head -1 gpurun_out/shapes_final_$wl.txt; done
tail -c 300 gpurun_out/bench_final.json
