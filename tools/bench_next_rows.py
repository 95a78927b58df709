"""Timing of the rows next to the denoising path: VAE decode and the embedding producers, ours vs the same oracle modules run
eagerly in 16-bit on the same GPU (stock PyTorch kernels).  CUDA events, median of a few repetitions.  One JSON line per item.
    python tools/bench_next_rows.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import embed_ref
from oracle.vae_ref import build_ref_vae, sd15_vae_config, sdxl_vae_config
from consistentid_b200.embed import FacialEncoder, ProjPlusModel
from consistentid_b200.vae import B200VAEDecoder

dev = "cuda"


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def init(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return {n: (torch.randn(s, generator=g) * (s[-1] ** -0.5) if len(s) > 1 else (1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(s, generator=g))
            for n, s in shapes.items()}


VAE_CASES = () if os.environ.get("SKIP_VAE") else (("vae_decode sd15 512x512 batch 8 fp16", sd15_vae_config(), 8, 64, torch.float16),
                            ("vae_decode sdxl 1024x1024 batch 4 bf16", sdxl_vae_config(), 4, 128, torch.bfloat16))
for name, cfg, B, h, dt in VAE_CASES:
    try:
        ref16 = build_ref_vae(cfg, dtype=dt).to(dev)
        eng = B200VAEDecoder(ref16.state_dict(), scaling_factor=cfg.scaling_factor, dtype=dt)
        z = (torch.randn(B, 4, h, h, device=dev) * cfg.scaling_factor).to(dt)
        with torch.no_grad():
            ms_ours = timeit(lambda: eng.decode_latents(z))
            ms_eager = timeit(lambda: ref16.decode_latents(z), reps=2)
            err = (eng.decode_latents(z).float() - ref16.decode_latents(z).float()).abs().max().item()
        print(json.dumps({"item": name, "ms_ours": round(ms_ours, 2), "ms_eager16_gpu": round(ms_eager, 2), "speedup": round(ms_eager / ms_ours, 2),
                          "images_per_s_ours": round(B / ms_ours * 1e3, 2), "max_abs_diff_vs_eager16": round(err, 4)}), flush=True)
        del ref16, eng, z
        torch.cuda.empty_cache()
    except Exception as e:
        print(json.dumps({"item": name, "error": repr(e)}), flush=True)

for name, cad, dt in (("embedding producers sd15 (ProjPlusModel + FacialEncoder, 1 identity, cond+uncond)", 768, torch.float16),
                      ("embedding producers sdxl", 2048, torch.bfloat16)):
    try:
        pm = ProjPlusModel(cross_attention_dim=cad, dtype=dt); sd_pm = init(pm.w._shapes, 1); pm.load_state_dict(sd_pm)
        fe = FacialEncoder(output_dim=cad, embed_dim=cad, dtype=dt); sd_fe = init(fe.w._shapes, 2); fe.load_state_dict(sd_fe)
        sd_pm16 = {k: v.to(dev, dt) for k, v in sd_pm.items()}; sd_fe16 = {k: v.to(dev, dt) for k, v in sd_fe.items()}
        idv = torch.randn(2, 512, device=dev).to(dt); clip = torch.randn(2, 257, 1280, device=dev).to(dt)
        prompt = torch.randn(2, 77, cad, device=dev).to(dt); imgs = torch.randn(2, 5, 257, 1280, device=dev).to(dt)
        cm = torch.zeros(2, 77, dtype=torch.bool, device=dev); vm = torch.zeros(2, 5, dtype=torch.bool, device=dev)
        cm[:, [3, 9, 20]] = True; vm[:, :3] = True

        def ours():
            pm(idv, clip); fe(prompt, imgs, cm, vm)

        def eager():
            embed_ref.proj_plus_model(sd_pm16, idv, clip); embed_ref.facial_encoder(sd_fe16, prompt, imgs, cm, vm)

        with torch.no_grad():
            ms_ours, ms_eager = timeit(ours, 5), timeit(eager, 5)
        print(json.dumps({"item": name, "ms_ours": round(ms_ours, 3), "ms_eager16_gpu": round(ms_eager, 3), "speedup": round(ms_eager / ms_ours, 2)}), flush=True)
    except Exception as e:
        print(json.dumps({"item": name, "error": repr(e)}), flush=True)
