#!/bin/bash
# One consolidated GPU validation pass (every command bounded by `timeout`): tests, smoke, benches, ncu launch lists.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1 )
( timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_sd15.json 2> gpurun_out/final_bench_sd15.err )
( timeout 600 python bench.py --workload sdxl --steps 3 --warmup 3 --no-cpu > gpurun_out/final_bench_sdxl.json 2> gpurun_out/final_bench_sdxl.err )
( timeout 600 python bench.py --workload sd15_cn --steps 3 --warmup 3 --no-cpu > gpurun_out/final_bench_sd15_cn.json 2> gpurun_out/final_bench_sd15_cn.err )
for wl in sd15 sdxl; do
  timeout 600 ncu --kernel-name-base demangled -k regex:cid:: --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches_$wl.csv python tools/profile_step.py $wl 2 > gpurun_out/final_prof_$wl.log 2>&1
  timeout 600 ncu --kernel-name-base demangled -k regex:cid:: --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/final_dram_$wl.csv python tools/profile_step.py $wl 1 > gpurun_out/final_dram_$wl.log 2>&1
  timeout 600 python tools/profile_shapes.py $wl > gpurun_out/final_shapes_$wl.txt 2>&1
done
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/final_smoke.log | tail -1; for f in sd15 sdxl sd15_cn; do cut -c1-110 gpurun_out/final_bench_$f.json; done
