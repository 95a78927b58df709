"""Whole-path parity at the BASELINE configs' REAL sizes (VERDICT r1 "next" #3): the full SD1.5 UNet (64^2 latent, head dims 40/80/160
composed) and the full SDXL UNet (128^2 latent, 10-layer transformer stacks), one forward at UNet batch 2, and a 30-step SD1.5 denoising
loop at B=1 - against the oracle evaluated in fp32 ON THE GPU (TF32 off; seconds instead of the CPU's minutes) with the oracle in eager
16-bit beside it.

Each case asserts the relative criterion of tests/test_unet_gpu.py AND reports, in north_star's own units, the fraction of output elements
inside rtol = atol = 1e-3 of the fp32 truth for ours and for eager-16.  north_star's "within rtol=1e-3/atol=1e-3 on the final latent" is
not met by ANY 16-bit evaluation of this network (the eager-16 column shows what the reference's own fp16 pipeline achieves against fp32);
the assertion therefore is "ours is at least as close to fp32 as eager-16 is, up to a small slack".  The numbers are appended to
gpurun_out/parity_fullsize.jsonl (copied to profiles/ by the builder).
"""
import copy
import json
import os

import pytest
import torch

from oracle import synth
from oracle.loop_ref import denoise_sd15
from oracle.schedulers_ref import make_scheduler
from oracle.unet_ref import sd15_config, sdxl_config
from tests.test_unet_gpu import ERR_FACTOR, ABS_FLOOR, _engine_from_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = ATOL = 1e-3           # north_star's tolerance on the final latent


def _inside(x, truth):
    x, truth = x.float().cpu(), truth.float().cpu()
    return float(((x - truth).abs() <= ATOL + RTOL * truth.abs()).float().mean())


def _report(name, ours, truth, eager, extra=None, rms_factor=1.5, frac_slack=0.02):
    ours, truth, eager = ours.float().cpu(), truth.float().cpu(), eager.float().cpu()
    scale = truth.abs().max().item()
    rec = {"case": name, "max_abs_truth": scale,
           "ours_max_abs_err": (ours - truth).abs().max().item(), "eager16_max_abs_err": (eager - truth).abs().max().item(),
           "ours_rms_err": (ours - truth).pow(2).mean().sqrt().item(), "eager16_rms_err": (eager - truth).pow(2).mean().sqrt().item(),
           "ours_frac_inside_rtol_atol_1e-3": _inside(ours, truth), "eager16_frac_inside_rtol_atol_1e-3": _inside(eager, truth)}
    if extra:
        rec.update(extra)
    print("[fullsize parity] " + json.dumps(rec))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_fullsize.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    assert torch.isfinite(ours).all()
    assert rec["ours_max_abs_err"] <= ERR_FACTOR * rec["eager16_max_abs_err"] + ABS_FLOOR * max(scale, 1.0), rec
    # rms error: ours no worse than 1.5 x eager-16 (it is measured BETTER: fp32 accumulation across the fused epilogues)
    assert rec["ours_rms_err"] <= rms_factor * rec["eager16_rms_err"] + 1e-4 * max(scale, 1.0), rec
    # north_star units: at least as many elements inside rtol/atol 1e-3 as the eager 16-bit evaluation of the reference, minus 2 points
    assert rec["ours_frac_inside_rtol_atol_1e-3"] >= rec["eager16_frac_inside_rtol_atol_1e-3"] - frac_slack, rec
    return rec


def _fp32_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _to_gpu(unet, dtype=None):
    u = unet.cuda() if dtype is None else unet.to(dtype).cuda()
    for p in u.attn_processors.values():
        p.cuda() if dtype is None else p.to(dtype).cuda()
    return u


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dtype", [("sd15", torch.float16), ("sdxl", torch.bfloat16)])
def test_full_unet_forward_parity(kind, dtype):
    _fp32_math()
    cfg = sd15_config() if kind == "sd15" else sdxl_config()
    ref = synth.build_ref_unet(cfg, rank=128)                      # fp32, architecture-exact, seeded
    NB, h = 2, cfg.sample_size
    null, aug, _ = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(NB, h, h, seed=3)
    ehs = torch.cat([null, aug])
    added = None
    if kind == "sdxl":
        g = torch.Generator().manual_seed(9)
        added = {"text_embeds": torch.randn(NB, 1280, generator=g), "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).expand(NB, -1).contiguous()}
    t = torch.tensor(601)
    eng = _engine_from_oracle(ref, dtype, 128)                      # packs the fp32 weights into the 16-bit arena (LoRA folded in fp32)
    with torch.no_grad():
        ref = _to_gpu(ref)
        added32 = None if added is None else {k: v.cuda() for k, v in added.items()}
        truth = ref(x.cuda(), t.cuda(), ehs.cuda(), added_cond_kwargs=added32).sample.cpu()
        ref16 = _to_gpu(ref, dtype)                                  # in place: the fp32 copy is no longer needed
        added16 = None if added is None else {k: v.cuda().to(dtype if k == "text_embeds" else v.dtype) for k, v in added.items()}
        eager = ref16(x.cuda().to(dtype), t.cuda(), ehs.cuda().to(dtype), added_cond_kwargs=added16).sample.cpu()
        del ref, ref16
        torch.cuda.empty_cache()
    out = eng(x.cuda().to(dtype), t, ehs.cuda().to(dtype), cross_attention_kwargs={}, added_cond_kwargs=added16).sample
    torch.cuda.synchronize()
    _report(f"unet_forward full {kind} {str(dtype)[6:]} NB={NB} latent {h}x{h}", out, truth, eager)


@pytest.mark.gpu
def test_full_sd15_denoise_loop_parity():
    """configs[1]'s loop at B = 1: final latents after 1, 5 and 30 DDIM steps (each a fresh run from the same noise)."""
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    _fp32_math()
    dtype = torch.float16
    cfg = sd15_config()
    ref = synth.build_ref_unet(cfg, rank=128)
    eng = _engine_from_oracle(ref, dtype, 128)
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    h = cfg.sample_size
    lat = synth.synth_latents(1, h, h, seed=0)
    den = B200Denoiser(eng, B200Scheduler("ddim"), use_cuda_graph=True)
    ref = _to_gpu(ref)
    ref16 = _to_gpu(copy.deepcopy(ref), dtype)
    c32, c16 = (lambda v: v.cuda()), (lambda v: v.cuda().to(dtype))
    for steps in (1, 5, 30):
        truth = denoise_sd15(ref, make_scheduler("ddim"), c32(lat), c32(null), c32(aug), c32(txt), steps, guidance_scale=5.0, start_merge_step=0).cpu()
        eager = denoise_sd15(ref16, make_scheduler("ddim"), c16(lat), c16(null), c16(aug), c16(txt), steps, guidance_scale=5.0, start_merge_step=0).cpu()
        out = den(lat, null, aug, txt, num_inference_steps=steps, guidance_scale=5.0, start_merge_step=0)
        torch.cuda.synchronize()
        # a 30-step loop amplifies rounding differences chaotically: either 16-bit evaluation can land closer to fp32 on a given seed
        _report(f"loop full sd15 fp16 B=1 {steps} ddim steps", out, truth, eager, {"steps": steps}, rms_factor=2.0, frac_slack=0.05)
