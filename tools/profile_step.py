"""Run N eager (non-graph) denoising iterations of a BASELINE workload - the command profiled with
`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_<wl>.csv python tools/profile_step.py <wl>`
(the per-launch list whose per-kernel SHARES back the roofline numbers of bench.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from consistentid_b200.arch import sd15_spec, sdxl_spec, synth_state_dicts
from consistentid_b200.pipeline import B200Denoiser
from consistentid_b200.scheduler import B200Scheduler
from consistentid_b200.unet import B200UNet

wl_name = sys.argv[1] if len(sys.argv) > 1 else "sd15"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[wl_name]
dtype = torch.float16 if wl["dtype"] == "fp16" else torch.bfloat16
spec = sd15_spec() if wl["model"] == "sd15" else sdxl_spec()
dev = torch.device("cuda")
usd, asd = synth_state_dicts(spec, dev, dtype)
unet = B200UNet(spec, usd, asd, dtype=dtype, device=dev)
del usd, asd
sched = B200Scheduler(wl["scheduler"])
den = B200Denoiser(unet, sched, use_cuda_graph=False)
B, h = wl["batch"], wl["res"] // 8
lat, prompts, extra = bench.synth_inputs(spec, B, h, h, 0, spec.addition_embed_type == "text_time")
out = den(lat.to(dev), prompts[0].to(dev), prompts[1].to(dev), prompts[2].to(dev), num_inference_steps=iters,
          guidance_scale=wl["guidance"], start_merge_step=-1, **{k: v.to(dev) for k, v in extra.items()})
torch.cuda.synchronize()
print("done", out.float().abs().mean().item())
