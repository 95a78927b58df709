#!/bin/bash
mkdir -p gpurun_out
CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace.so timeout 120 python tools/trace_attn.py sd15 > gpurun_out/trace_attn6b_sd15.txt 2>&1; tail -8 gpurun_out/trace_attn6b_sd15.txt
CID_LIB_PATH=$PWD/tools/bin/libcidb200_trace_notoken.so timeout 120 python tools/trace_attn.py sd15 > gpurun_out/trace_attn6c_sd15.txt 2>&1; cat gpurun_out/trace_attn6c_sd15.txt
# GEMM epilogue A/B: 16 vs 8 epilogue warps
timeout 400 python -m pytest tests/test_kernels_gpu.py -k "gemm or conv" -x -q -p no:cacheprovider > gpurun_out/pytest_gemm16.log 2>&1; echo "gemm16 rc=$?"; tail -2 gpurun_out/pytest_gemm16.log
for wl in sd15 sdxl; do
  timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes6_epi16_$wl.txt 2>&1
  CID_LIB_PATH=$PWD/tools/bin/libcidb200_epi8.so timeout 300 python tools/profile_shapes.py $wl > gpurun_out/shapes6_epi8_$wl.txt 2>&1
  head -1 gpurun_out/shapes6_epi16_$wl.txt; head -1 gpurun_out/shapes6_epi8_$wl.txt
done
python - <<'P'
import re
for wl in ("sd15","sdxl"):
    def load(f):
        d={}
        for l in open(f):
            p=l.split()
            if len(p)>8 and p[0] in ("gemm","conv3x3"): d[tuple(p[:5])]=(int(p[5]),float(p[6]),float(p[7]))
        return d
    a,b=load(f"gpurun_out/shapes6_epi16_{wl}.txt"),load(f"gpurun_out/shapes6_epi8_{wl}.txt")
    ta=sum(v[1] for v in a.values()); tb=sum(v[1] for v in b.values())
    print(wl,"gemm+conv ms/iter: epi16",round(ta,3),"epi8",round(tb,3))
    for k in sorted(a,key=lambda k:-b.get(k,(0,0,0))[1])[:14]:
        print("  ",k,a[k][1],"vs",b.get(k,(0,0,0))[1])
P
