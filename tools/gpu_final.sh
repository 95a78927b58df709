#!/bin/bash
# Final validation + evidence run (one GPU): whole -m gpu suite, smoke, driver-style bench + reference arm, ncu launch lists of one eager
# denoising iteration (OUR kernels only) and full captures of the top kernels.  Only text extracts are kept (gpurun_out/ must stay < 64 MiB).
mkdir -p gpurun_out
KREG='regex:gemm_tc2|attn_|gn_|layernorm|upsample|phase_split|cfg_sched|skinny|pack_cross|latents_to|advance_step|timestep_embed|nchw|rows_to'
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 1500 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; echo "ref rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for wl in sd15 sdxl; do
  timeout 600 ncu --metrics $M --clock-control none -k "$KREG" --csv --log-file gpurun_out/ncu_step_$wl.csv python tools/profile_step.py $wl 1 > gpurun_out/ncu_step_$wl.log 2>&1
  python tools/summarize_ncu_step.py gpurun_out/ncu_step_$wl.csv > gpurun_out/ncu_step_${wl}_summary.txt 2>&1; head -16 gpurun_out/ncu_step_${wl}_summary.txt
  python tools/summarize_dram.py gpurun_out/ncu_step_$wl.csv gpurun_out/dram_traffic_$wl.json > gpurun_out/dram_${wl}_summary.txt 2>&1
  gzip -f gpurun_out/ncu_step_$wl.csv
done
cap() {  # name, kernel regex, model, case
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o /tmp/ncu_$1 python tools/profile_kernels.py $3 $4 > /dev/null 2>&1
  python tools/ncu_metrics.py /tmp/ncu_$1.ncu-rep > gpurun_out/ncu_$1_metrics.txt 2>&1
  python tools/ncu_hot.py /tmp/ncu_$1.ncu-rep 30 > gpurun_out/ncu_$1_hot.txt 2>&1
  grep -E "duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|pipe_xu.avg|dram__bytes_read.sum =" gpurun_out/ncu_$1_metrics.txt | head -6
}
cap attn7_sd15 attn_self7 sd15 attn_self
cap attn7_sdxl attn_self7 sdxl attn_self
cap conv_sd15 gemm_tc2 sd15 conv3x3
cap ff1_sdxl gemm_tc2 sdxl gemm_ff1
cap outproj_sd15 gemm_tc2 sd15 gemm_out_proj
cap cross2_sd15 attn_cross2 sd15 attn_cross
cap cross2_sdxl attn_cross2 sdxl attn_cross
du -sh gpurun_out
tail -c 400 gpurun_out/bench_final.json
