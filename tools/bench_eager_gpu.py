"""the "reference diffusers-CUDA pipeline" stand-in - the CPU oracle's modules (reference attention processors inside the
restated diffusers UNet, LoRA unfolded, naive [2B*H,N,N] self-attention exactly as attention.py:157-158 without xformers) run EAGERLY in
16-bit on one B200 with stock PyTorch kernels (cuBLAS / cuDNN), no CUDA graph, no fusion.  Reported next to bench.py's numbers; it is a
measurement tool, not part of the product path.
    python tools/bench_eager_gpu.py sd15|sdxl [denoise_steps_timed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import synth
from oracle.loop_ref import denoise_sd15, denoise_sdxl
from oracle.schedulers_ref import make_scheduler
from oracle.unet_ref import sd15_config, sdxl_config

wl_name = sys.argv[1] if len(sys.argv) > 1 else "sd15"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wl = bench.WORKLOADS[wl_name]
dtype = torch.float16 if wl["dtype"] == "fp16" else torch.bfloat16
cfg = sd15_config() if wl["model"] == "sd15" else sdxl_config()
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
unet = synth.build_ref_unet(cfg, rank=128, dtype=dtype).to(dev)
for p in unet.attn_processors.values():
    p.to(dev)
B, h = wl["batch"], wl["res"] // 8
cad = cfg.cross_attention_dim
null, aug, txt = (t.to(dev, dtype) for t in synth.synth_prompts(cad))
sched = make_scheduler(wl["scheduler"])
sched.set_timesteps(n_steps)
lat = synth.synth_latents(B, h, h, seed=0, init_noise_sigma=float(sched.init_noise_sigma)).to(dev, dtype)


def run(steps):
    s = make_scheduler(wl["scheduler"])
    if wl["model"] == "sd15":
        return denoise_sd15(unet, s, lat, null, aug, txt, steps, guidance_scale=wl["guidance"], start_merge_step=0)
    g = torch.Generator().manual_seed(5)
    pooled = [torch.randn(1, 1280, generator=g).to(dev, dtype) for _ in range(3)]
    tid = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]], device=dev, dtype=dtype)
    return denoise_sdxl(unet, s, lat, null, txt, null, aug, pooled[0], pooled[1], pooled[2], tid, steps, guidance_scale=wl["guidance"], start_merge_step=0)


run(2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = run(n_steps); e1.record(); torch.cuda.synchronize()
ms_iter = e0.elapsed_time(e1) / n_steps
res = {"impl": "eager-gpu oracle (torch 16-bit, stock kernels, no graph)", "workload": wl_name, "per_gpu_batch": B, "dtype": wl["dtype"],
       "ms_per_denoise_iteration": round(ms_iter, 2), "images_per_sec_at_%d_steps" % wl["denoise_steps"]: round(B / (ms_iter * wl["denoise_steps"] / 1e3), 4),
       "iterations_timed": n_steps, "finite": bool(torch.isfinite(out.float()).all()), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
print(json.dumps(res), flush=True)
