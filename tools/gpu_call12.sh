#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_kernels12.log 2>&1; rc=$?; echo "kernels rc=$rc"; tail -8 gpurun_out/pytest_kernels12.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_vae_gpu.py tests/test_clip_gpu.py tests/test_embed_gpu.py tests/test_checkpoint_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet12.log 2>&1; echo "unet rc=$?"; tail -4 gpurun_out/pytest_unet12.log
for wl in sd15 sdxl; do
  timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes12_$wl.txt 2>&1
  CID_LIB_PATH=$PWD/tools/bin/libcidb200_notma.so timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes12_notma_$wl.txt 2>&1
  head -1 gpurun_out/shapes12_$wl.txt; head -1 gpurun_out/shapes12_notma_$wl.txt
done
python - <<'P'
for wl in ("sd15","sdxl"):
    def load(f):
        d={}
        for l in open(f):
            p=l.split()
            if len(p)>8 and p[0] in ("gemm","conv3x3"): d[tuple(p[:5])]=(int(p[5]),float(p[6]),float(p[7]))
        return d
    a,b=load(f"gpurun_out/shapes12_{wl}.txt"),load(f"gpurun_out/shapes12_notma_{wl}.txt")
    print(wl,"gemm+conv ms/iter: tma",round(sum(v[1] for v in a.values()),3),"register epilogue",round(sum(v[1] for v in b.values()),3))
    for k in sorted(a,key=lambda k:-b.get(k,(0,0,0))[1])[:14]:
        print("  ",k,a[k][1],a[k][2],"vs",b.get(k,(0,0,0))[1])
P
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench12_$wl.json 2> gpurun_out/bench12_$wl.err; python -c "
import json;d=json.loads(open('gpurun_out/bench12_$wl.json').read().strip().splitlines()[-1]);print('$wl',d['value'],d['ms_per_step'],d['clocks'])"; done
