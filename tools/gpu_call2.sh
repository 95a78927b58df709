#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_vae_gpu.py tests/test_fullsize_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/pytest_unet.log
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager > gpurun_out/bench2_$wl.json 2> gpurun_out/bench2_$wl.err; done
for wl in sd15 sdxl; do timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes2_$wl.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_self4 -s 3 -c 1 -o gpurun_out/ncu_attn4_sd15 python tools/profile_kernels.py sd15 attn_self > gpurun_out/ncu_attn4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o gpurun_out/ncu_outproj_sd15 python tools/profile_kernels.py sd15 gemm_out_proj > gpurun_out/ncu_outproj.log 2>&1
python - <<'P'
import json
for wl in ("sd15","sdxl"):
    try:
        d=json.loads(open(f"gpurun_out/bench2_{wl}.json").read().strip().splitlines()[-1])
        print(wl, d["value"], d["ms_per_step"], {k:(v["ms"]) for k,v in d.get("kernels",{}).items()})
    except Exception as e: print(wl, "ERR", e)
P
