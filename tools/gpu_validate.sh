#!/bin/bash
# One-call GPU validation (run under gpurun from the repo root): microbenchmarks, the whole -m gpu suite, the driver-style bench line and
# the per-shape tables.  Every step has its own timeout; outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_start.csv 2>&1
[ -x tools/bin/mb_softmax ] && timeout 120 tools/bin/mb_softmax > gpurun_out/mb_softmax.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err; echo "bench rc=$?"
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes_$wl.txt 2>&1; done
tail -c 1500 gpurun_out/bench_all.json
