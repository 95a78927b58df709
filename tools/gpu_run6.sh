#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1 )
( timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_sd15.json 2> gpurun_out/final_bench_sd15.err )
( timeout 600 python bench.py --workload sdxl --steps 3 --warmup 3 --no-cpu > gpurun_out/final_bench_sdxl.json 2> gpurun_out/final_bench_sdxl.err )
( timeout 600 python bench.py --workload sd15_cn --steps 3 --warmup 3 --no-cpu > gpurun_out/final_bench_sd15_cn.json 2> gpurun_out/final_bench_sd15_cn.err )
for wl in sd15 sdxl; do timeout 600 python tools/bench_eager_gpu.py $wl 4 > gpurun_out/eager_gpu_$wl.json 2> gpurun_out/eager_gpu_$wl.err; done
for wl in sd15 sdxl; do
  timeout 600 ncu --kernel-name-base demangled -k regex:cid:: --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches_$wl.csv python tools/profile_step.py $wl 2 > gpurun_out/final_prof_$wl.log 2>&1
done
tail -3 gpurun_out/final_pytest.log; tail -1 gpurun_out/final_smoke.log
for f in sd15 sdxl sd15_cn; do cut -c1-110 gpurun_out/final_bench_$f.json; done
cat gpurun_out/eager_gpu_sd15.json gpurun_out/eager_gpu_sdxl.json; tail -3 gpurun_out/eager_gpu_sdxl.err
