"""Per-shape time of the tensor-core launches of ONE eager denoising iteration (CUDA events per launch):
    python tools/profile_shapes.py sd15|sdxl
prints, per (kind, M, N, K): launches, total ms, TFLOP/s, tiles of the N-tile the library picks and the wave efficiency on 148 SMs."""
import os, sys, math, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from consistentid_b200 import lib
from consistentid_b200.arch import sd15_spec, sdxl_spec, synth_state_dicts
from consistentid_b200.pipeline import B200Denoiser
from consistentid_b200.scheduler import B200Scheduler
from consistentid_b200.unet import B200UNet

if os.environ.get("CID_TOOL_SPLITK"):          # tail-balancing policy override for experiments: "max_split,min_kblocks"
    lib.set_splitk(*[int(v) for v in os.environ["CID_TOOL_SPLITK"].split(",")])
wl_name = sys.argv[1] if len(sys.argv) > 1 else "sd15"
wl = bench.WORKLOADS[wl_name]
dtype = torch.float16 if wl["dtype"] == "fp16" else torch.bfloat16
spec = sd15_spec() if wl["model"] == "sd15" else sdxl_spec()
dev = torch.device("cuda")
usd, asd = synth_state_dicts(spec, dev, dtype)
unet = B200UNet(spec, usd, asd, dtype=dtype, device=dev)
del usd, asd
den = B200Denoiser(unet, B200Scheduler(wl["scheduler"]), use_cuda_graph=False)
B, h = wl["batch"], wl["res"] // 8
lat, prompts, extra = bench.synth_inputs(spec, B, h, h, 0, spec.addition_embed_type == "text_time")
kw = dict(guidance_scale=wl["guidance"], start_merge_step=-1, **{k: v.to(dev) for k, v in extra.items()})
args = (lat.to(dev), prompts[0].to(dev), prompts[1].to(dev), prompts[2].to(dev))
den(*args, num_inference_steps=2, **kw)
agg = collections.OrderedDict()
REPS = 3
for rep in range(REPS):
    den(*args, num_inference_steps=1, profile=True, **kw)
    for r in den.last_profile:
        a = agg.setdefault((r["kind"], r["shape"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += r["ms"]; a[2] += r["flops"]; a[3] += r["bytes"]
tot = sum(a[1] for a in agg.values()) / REPS
print(f"{wl_name}: profiled launches of one iteration: {tot:.3f} ms (CUDA events per launch, eager, no PDL overlap)")
print(f"{'kind':14s} {'M':>6s} {'N':>6s} {'K':>6s} epi   n   ms/iter   TF/s    GB/s  tiles waves eff")
for (kind, sh), (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    n //= REPS; ms /= REPS; fl /= REPS; by /= REPS
    tf = fl / ms / 1e9 if fl else 0.0
    gbs = by / ms / 1e6
    if kind not in ("gemm", "conv3x3"):
        desc = "" if sh is None else "x".join(str(v) for v in sh)
        print(f"{kind:14s} {desc:27s} {n:3d} {ms:8.3f} {tf:7.1f} {gbs:7.0f}")
        continue
    M, N, K, epi = sh
    bn = lib.gemm_tile_n(N, epi)
    tiles = math.ceil(M / 128) * math.ceil(N / bn)
    waves = tiles / 148
    print(f"{kind:14s} {M:6d} {N:6d} {K:6d} {epi:3d} {n:3d} {ms:8.3f} {tf:7.1f} {gbs:7.0f} {tiles:6d} {waves:5.2f} {waves / math.ceil(waves):4.2f}  BN={bn}")
