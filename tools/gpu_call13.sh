#!/bin/bash
# call 13: validate attn_cross2 + the wave-model tile choice; A/B against tools/bin/libcidb200_before.so (static tiles, attn_cross v1)
mkdir -p gpurun_out
# the new kernel first, alone and under a short timeout: if it fails or hangs, the rest of the call runs with the v1 cross-attention build
timeout 180 python -m pytest tests/test_kernels_gpu.py -k attn_cross -q -p no:cacheprovider > gpurun_out/pytest_cross13.log 2>&1; rc=$?; echo "cross rc=$rc"; tail -12 gpurun_out/pytest_cross13.log
if [ $rc -ne 0 ]; then export CID_LIB_PATH=$PWD/tools/bin/libcidb200_crossv1.so; echo "FALLING BACK to $CID_LIB_PATH"; fi
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_kernels13.log 2>&1; rc=$?; echo "kernels rc=$rc"; tail -12 gpurun_out/pytest_kernels13.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_controlnet_gpu.py tests/test_processors_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_unet13.log 2>&1; echo "unet rc=$?"; tail -4 gpurun_out/pytest_unet13.log
for wl in sd15 sdxl; do
  timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes13_$wl.txt 2>&1
  CID_LIB_PATH=$PWD/tools/bin/libcidb200_before.so timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes13_before_$wl.txt 2>&1
  head -1 gpurun_out/shapes13_$wl.txt; head -1 gpurun_out/shapes13_before_$wl.txt
done
python - <<'P'
for wl in ("sd15","sdxl"):
    def load(f):
        d={}
        for l in open(f):
            p=l.split()
            if len(p)>8 and p[0] in ("gemm","conv3x3"): d[tuple(p[:5])]=(int(p[5]),float(p[6]),float(p[7]),p[-1])
            elif len(p)>4 and p[0].startswith("attn_"): d[tuple(p[:2])]=(int(p[2]),float(p[3]),float(p[4]),"")
        return d
    a,b=load(f"gpurun_out/shapes13_{wl}.txt"),load(f"gpurun_out/shapes13_before_{wl}.txt")
    print(wl,"gemm+conv+attn ms/iter: new",round(sum(v[1] for v in a.values()),3),"before",round(sum(v[1] for v in b.values()),3))
    for k in sorted(a,key=lambda k:-abs(a[k][1]-b.get(k,(0,0,0,""))[1]))[:16]:
        print("  ",k,a[k][1],a[k][2],a[k][3],"vs",b.get(k,(0,0,0,""))[1],b.get(k,(0,0,0,""))[3])
P
for wl in sd15 sdxl; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-eager --no-profile > gpurun_out/bench13_$wl.json 2> gpurun_out/bench13_$wl.err; python -c "
import json;d=json.loads(open('gpurun_out/bench13_$wl.json').read().strip().splitlines()[-1]);print('$wl',d['value'],d['ms_per_step'],d['clocks'])"; done
