#!/bin/bash
# One-call GPU validation (run under gpurun from the repo root): microbenchmarks, the whole -m gpu suite, the driver-style bench line and
# the per-shape tables.  Every step has its own timeout; outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_start.csv 2>&1
[ -x tools/bin/mb_softmax ] && timeout 120 tools/bin/mb_softmax > gpurun_out/mb_softmax.txt 2>&1
# new attention kernel first, on its own short leash: if it fails or hangs the rest of the run uses the round-1 kernel (A/B variant library)
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attn_self" -x -q -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1
rc=$?; echo "attn rc=$rc" >> gpurun_out/pytest_attn.log; tail -3 gpurun_out/pytest_attn.log
if [ $rc -ne 0 ] && [ -f tools/bin/libcidb200_v3.so ]; then export CID_LIB_PATH=$PWD/tools/bin/libcidb200_v3.so; echo "FALLING BACK TO $CID_LIB_PATH"; fi
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err; echo "bench rc=$?"
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes_$wl.txt 2>&1; done
if [ -z "$CID_LIB_PATH" ] && [ -f tools/bin/libcidb200_v3.so ]; then
  for wl in sd15 sdxl; do CID_LIB_PATH=$PWD/tools/bin/libcidb200_v3.so timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes_${wl}_attnv3.txt 2>&1; done
fi
grep -h attn_self gpurun_out/shapes_*.txt | head -20
tail -c 1500 gpurun_out/bench_all.json
