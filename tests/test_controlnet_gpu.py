"""GPU parity of config 5 (ControlNet + inpaint, pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:375-449):
B200ControlNet forward and the fused ControlNet -> UNet -> CFG -> step -> blend loop vs the CPU oracle."""
import pytest
import torch

from oracle import synth
from oracle.controlnet_ref import build_ref_controlnet
from oracle.loop_ref import denoise_controlnet_inpaint
from oracle.schedulers_ref import make_scheduler
from oracle.unet_ref import tiny_config
from tests.test_unet_gpu import _cmp, _engine_from_oracle

COND_CH = (16, 32, 64, 64)


def _cn_engine(cn_ref, dtype):
    from consistentid_b200.arch import UNetSpec
    from consistentid_b200.controlnet import B200ControlNet
    return B200ControlNet(UNetSpec.from_config(cn_ref.config), cn_ref.state_dict(), dtype=dtype, device="cuda", cond_block_out_channels=COND_CH)


@pytest.mark.gpu
def test_controlnet_forward_parity():
    dtype = torch.float16
    cfg = tiny_config("sd15")
    cn = build_ref_controlnet(cfg, cond_block_out_channels=COND_CH)
    B, h = 2, cfg.sample_size
    _, aug, _ = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(B, h, h, seed=5)
    ehs = aug.expand(B, -1, -1).contiguous()
    ctrl = torch.rand(B, 3, 8 * h, 8 * h, generator=torch.Generator().manual_seed(3))
    t = torch.tensor(401)
    with torch.no_grad():
        down_t, mid_t = cn(x, t, ehs, ctrl, conditioning_scale=0.5)
        cn16 = build_ref_controlnet(cfg, dtype=dtype, cond_block_out_channels=COND_CH).cuda()
        down_e, mid_e = cn16(x.cuda().to(dtype), t.cuda(), ehs.cuda().to(dtype), ctrl.cuda().to(dtype), conditioning_scale=0.5)
    eng = _cn_engine(cn, dtype)
    down_o, mid_o = eng(x.cuda().to(dtype), t, ehs.cuda().to(dtype), ctrl.cuda(), conditioning_scale=0.5)
    torch.cuda.synchronize()
    assert len(down_o) == len(down_t) == 12
    for j, (o, tr, e) in enumerate(zip(list(down_o) + [mid_o], list(down_t) + [mid_t], list(down_e) + [mid_e])):
        assert o.shape == tr.shape
        _cmp(f"controlnet residual {j}", o, tr, e)


@pytest.mark.gpu
@pytest.mark.parametrize("nine", [False, True])
def test_controlnet_inpaint_loop(nine):
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    dtype = torch.float16
    cfg = tiny_config("sd15")
    if nine:
        cfg.in_channels = 9
    ref = synth.build_ref_unet(cfg, rank=16)
    cn_cfg = tiny_config("sd15")
    cn = build_ref_controlnet(cn_cfg, cond_block_out_channels=COND_CH)
    steps, B, h = 4, 2, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    g = torch.Generator().manual_seed(11)
    lat = synth.synth_latents(B, h, h, seed=0)
    img, noise = synth.synth_latents(B, h, h, seed=7), synth.synth_latents(B, h, h, seed=8)
    ctrl = torch.rand(B, 3, 8 * h, 8 * h, generator=g)
    mask = torch.zeros(B, 1, h, h)
    mask[:, :, h // 4: 3 * h // 4, h // 4: 3 * h // 4] = 1
    mil = img * (1 - mask) if nine else None
    kw = dict(guidance_scale=5.0, start_merge_step=1, conditioning_scale=0.5)
    truth = denoise_controlnet_inpaint(ref, cn, make_scheduler("ddim"), lat, null, aug, txt, ctrl, img, noise, mask, steps,
                                       masked_image_latents=mil, **kw)
    c = lambda t_: None if t_ is None else t_.cuda().to(dtype)
    ref16 = synth.build_ref_unet(cfg, rank=16, dtype=dtype).cuda()
    for p in ref16.attn_processors.values():
        p.cuda()
    cn16 = build_ref_controlnet(cn_cfg, dtype=dtype, cond_block_out_channels=COND_CH).cuda()
    eager = denoise_controlnet_inpaint(ref16, cn16, make_scheduler("ddim"), c(lat), c(null), c(aug), c(txt), c(ctrl), c(img), c(noise),
                                       c(mask), steps, masked_image_latents=c(mil), **kw)
    eng = _engine_from_oracle(ref, dtype, 16)
    cne = _cn_engine(cn, dtype)
    den = B200Denoiser(eng, B200Scheduler("ddim"), use_cuda_graph=True)
    out = den.controlnet_inpaint(cne, lat, null, aug, txt, ctrl, img, noise, mask, num_inference_steps=steps,
                                 masked_image_latents=mil, **kw)
    torch.cuda.synchronize()
    _cmp(f"controlnet+inpaint loop nine={nine}", out, truth, eager)


@pytest.mark.gpu
@pytest.mark.parametrize("nine", [False, True])
def test_plain_inpaint_loop(nine):
    """pipelines/StableDIffusionInpaint_ConsistentID.py:305-359: the inpaint loop WITHOUT a ControlNet (4-channel blend / 9-channel concat)."""
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    from oracle.loop_ref import denoise_inpaint
    dtype = torch.float16
    cfg = tiny_config("sd15")
    if nine:
        cfg.in_channels = 9
    ref = synth.build_ref_unet(cfg, rank=16)
    steps, B, h = 4, 2, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    lat = synth.synth_latents(B, h, h, seed=0)
    img, noise = synth.synth_latents(B, h, h, seed=7), synth.synth_latents(B, h, h, seed=8)
    mask = torch.zeros(B, 1, h, h)
    mask[:, :, h // 4: 3 * h // 4, h // 4: 3 * h // 4] = 1
    mil = img * (1 - mask) if nine else None
    kw = dict(guidance_scale=5.0, start_merge_step=1)
    truth = denoise_inpaint(ref, make_scheduler("ddim"), lat, null, aug, txt, img, noise, mask, steps, masked_image_latents=mil, **kw)
    c = lambda t_: None if t_ is None else t_.cuda().to(dtype)
    ref16 = synth.build_ref_unet(cfg, rank=16, dtype=dtype).cuda()
    for p in ref16.attn_processors.values():
        p.cuda()
    eager = denoise_inpaint(ref16, make_scheduler("ddim"), c(lat), c(null), c(aug), c(txt), c(img), c(noise), c(mask), steps,
                            masked_image_latents=c(mil), **kw)
    eng = _engine_from_oracle(ref, dtype, 16)
    den = B200Denoiser(eng, B200Scheduler("ddim"), use_cuda_graph=True)
    out = den.inpaint(lat, null, aug, txt, img, noise, mask, num_inference_steps=steps, masked_image_latents=mil, **kw)
    torch.cuda.synchronize()
    _cmp(f"plain inpaint loop nine={nine}", out, truth, eager)
