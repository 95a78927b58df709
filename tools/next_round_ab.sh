#!/bin/bash
# First GPU call of the next round: validate what was written after the round-1 GPU budget ran out, in one pass.
#   (here, before the call)  git apply tools/patches/gemm_v6_specialised_epilogue.patch tools/patches/clip_vision_encoder_ragged_attention.patch
#                            python -c "import __graft_entry__ as g; g.build()"
#   gpurun --timeout 1800 -- 'bash tools/next_round_ab.sh'
mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/ab.log; ( "$@" ) >> gpurun_out/ab.log 2>&1; echo "rc=$?" >> gpurun_out/ab.log; }
# 1. the whole suite with the defaults (includes test_clip_gpu.py and the ragged-attention checks once the CLIP patch is applied)
run timeout 1200 python -m pytest tests -m gpu -q
# 2. skinny_linear v2 as the default candidate: the suite again, then the producers' timing
CID_SKINNY_VERSION=2 run timeout 1200 python -m pytest tests -m gpu -q -k "skinny or time_embed or unet or denoise or embed or controlnet"
CID_SKINNY_VERSION=2 SKIP_VAE=1 run timeout 200 python tools/bench_next_rows.py
# 3. epilogue-specialised GEMM (CID_GEMM_VERSION=6, needs the v6 patch): kernel checks + UNet parity, then A/B
CID_GEMM_VERSION=6 run timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q
for v in 2 6; do for wl in sd15 sdxl; do
  CID_GEMM_VERSION=$v timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-profile > gpurun_out/ab_gemm${v}_$wl.json 2>> gpurun_out/ab.log
done; done
for m in sd15 sdxl; do CID_GEMM_VERSION=6 run timeout 300 python tools/profile_kernels.py $m gemm conv; done
grep -E "^== |^rc=|passed|failed|\"item\"" gpurun_out/ab.log | tail -40
for f in gpurun_out/ab_gemm*.json; do echo $f; cut -c1-100 $f; done
