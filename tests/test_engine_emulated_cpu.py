"""Host logic of the engines on the CPU: the real engine code (weight packing, LoRA folding, buffer plan, launch sequence, scheduler tables)
runs against tests/emulated_ops.py - torch restatements of the kernels - and must reproduce the oracle.  Complements the GPU parity tests:
a wrong offset, stride, weight layout or coefficient in the HOST code fails here without a GPU."""
import pytest
import torch

from oracle import synth
from oracle.loop_ref import denoise_sd15, denoise_sdxl
from oracle.schedulers_ref import make_scheduler
from oracle.unet_ref import tiny_config
from tests import emulated_ops


@pytest.fixture
def emu(monkeypatch):
    monkeypatch.setenv("CID_EMBED_GRAPH", "0")
    emulated_ops.install(monkeypatch)


def _close(name, got, want, tol):
    err = (got.float() - want.float()).abs().max().item()
    ref = want.abs().max().item()
    assert err <= tol * max(ref, 1.0), (name, err, ref)


def _engine(ref, dtype=torch.float32):
    from consistentid_b200.arch import UNetSpec
    from consistentid_b200.unet import B200UNet
    sd = {k: v for k, v in ref.state_dict().items() if ".processor." not in k}
    return B200UNet(UNetSpec.from_config(ref.config), sd, synth.adapter_state_dict(ref), dtype=dtype, device="cpu", rank=16)


@pytest.mark.parametrize("kind", ["sd15", "sdxl"])
def test_unet_forward_host_logic(emu, kind):
    cfg = tiny_config(kind)
    ref = synth.build_ref_unet(cfg, rank=16)
    B, h = 2, cfg.sample_size
    null, aug, _ = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(2 * B, h, h, seed=3)
    ehs = torch.cat([null.expand(B, -1, -1), aug.expand(B, -1, -1)])
    added = None
    if kind == "sdxl":
        g = torch.Generator().manual_seed(9)
        added = {"text_embeds": torch.randn(2 * B, cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim, generator=g),
                 "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).expand(2 * B, -1).contiguous()}
    t = torch.tensor(601)
    with torch.no_grad():
        want = ref(x, t, ehs, added_cond_kwargs=added).sample
    eng = _engine(ref)
    got = eng(x, t, ehs, cross_attention_kwargs={}, added_cond_kwargs=added).sample
    _close(f"unet forward {kind}", got, want, 2e-4)
    got2 = eng(x, t, ehs, cross_attention_kwargs={}, added_cond_kwargs=added).sample        # cached prompt K/V, reused buffers
    _close(f"unet forward {kind} (second call)", got2, want, 2e-4)


@pytest.mark.parametrize("sched_kind", ["ddim", "euler", "dpmpp2m"])
def test_denoise_loop_host_logic(emu, sched_kind):
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    cfg = tiny_config("sd15")
    ref = synth.build_ref_unet(cfg, rank=16)
    steps, B, h = 4, 2, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    sch = make_scheduler(sched_kind)
    sch.set_timesteps(steps)
    lat = synth.synth_latents(B, h, h, seed=0, init_noise_sigma=float(sch.init_noise_sigma))
    want = denoise_sd15(ref, sch, lat, null, aug, txt, steps, guidance_scale=5.0, start_merge_step=1)
    den = B200Denoiser(_engine(ref), B200Scheduler(sched_kind), use_cuda_graph=False)
    got = den(lat, null, aug, txt, num_inference_steps=steps, guidance_scale=5.0, start_merge_step=1)
    _close(f"loop sd15 {sched_kind}", got, want, 5e-4)


def test_denoise_loop_sdxl_host_logic(emu):
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    cfg = tiny_config("sdxl")
    ref = synth.build_ref_unet(cfg, rank=16)
    steps, B, h = 3, 1, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    g = torch.Generator().manual_seed(5)
    pdim = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    pooled = [torch.randn(1, pdim, generator=g) for _ in range(3)]
    tid = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]])
    sch = make_scheduler("euler")
    sch.set_timesteps(steps)
    lat = synth.synth_latents(B, h, h, seed=0, init_noise_sigma=float(sch.init_noise_sigma))
    want = denoise_sdxl(ref, sch, lat, null, txt, null, aug, pooled[0], pooled[1], pooled[2], tid, steps, guidance_scale=7.5, start_merge_step=0)
    den = B200Denoiser(_engine(ref), B200Scheduler("euler"), use_cuda_graph=False)
    got = den(lat, null, aug, txt, num_inference_steps=steps, guidance_scale=7.5, start_merge_step=0, neg_pooled=pooled[0],
              pooled_text_only=pooled[1], pooled_facial=pooled[2], add_time_ids=tid)
    _close("loop sdxl euler", got, want, 5e-4)


def test_vae_decode_host_logic(emu):
    from consistentid_b200.vae import B200VAEDecoder
    from oracle.vae_ref import build_ref_vae, tiny_vae_config
    cfg = tiny_vae_config()
    ref = build_ref_vae(cfg)
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1)) * cfg.scaling_factor * 3
    with torch.no_grad():
        want = ref.decode_latents(z)
    eng = B200VAEDecoder(ref.state_dict(), scaling_factor=cfg.scaling_factor, block_out_channels=cfg.block_out_channels, dtype=torch.float32, device="cpu")
    _close("vae decode", eng.decode_latents(z), want, 2e-4)


def test_embedding_producers_host_logic(emu):
    from consistentid_b200.embed import FacialEncoder, ProjPlusModel
    from oracle import embed_ref
    from tests.test_embed_gpu import _init
    dt = torch.float16                                           # the producers insist on 16-bit inputs; the emulation does fp32 math, 16-bit storage
    pm = ProjPlusModel(cross_attention_dim=128, id_embeddings_dim=64, clip_embeddings_dim=128, dtype=dt, device="cpu")
    sd = _init(pm.w._shapes, 31)
    pm.load_state_dict(sd)
    g = torch.Generator().manual_seed(32)
    idv, clip = torch.randn(2, 64, generator=g), torch.randn(2, 17, 128, generator=g)
    for shortcut in (False, True):
        want = embed_ref.proj_plus_model(sd, idv, clip, shortcut=shortcut, scale=0.7)
        _close("ProjPlusModel", pm(idv.to(dt), clip.to(dt), shortcut=shortcut, scale=0.7), want, 1e-2)
    fe = FacialEncoder(embedding_dim=128, output_dim=128, embed_dim=128, dtype=dt, device="cpu", dim=128, depth=2, heads=2)
    sd = _init(fe.w._shapes, 41)
    fe.load_state_dict(sd)
    prompt, imgs = torch.randn(2, 77, 128, generator=g), torch.randn(2, 5, 17, 128, generator=g)
    cm, vm = torch.zeros(2, 77, dtype=torch.bool), torch.zeros(2, 5, dtype=torch.bool)
    cm[0, [5, 9, 20]] = True; cm[1, [1, 76]] = True; vm[0, :3] = True; vm[1, :2] = True
    want = embed_ref.facial_encoder(sd, prompt, imgs, cm, vm)
    _close("FacialEncoder", fe(prompt.to(dt), imgs.to(dt), cm, vm), want, 1e-2)


COND_CH = (16, 32, 64, 64)


def _cn_engine(cn_ref):
    from consistentid_b200.arch import UNetSpec
    from consistentid_b200.controlnet import B200ControlNet
    return B200ControlNet(UNetSpec.from_config(cn_ref.config), cn_ref.state_dict(), dtype=torch.float32, device="cpu", cond_block_out_channels=COND_CH)


def test_controlnet_forward_host_logic(emu):
    from oracle.controlnet_ref import build_ref_controlnet
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
    down_o, mid_o = _cn_engine(cn)(x, t, ehs, ctrl, conditioning_scale=0.5)
    assert len(down_o) == len(down_t) == 12
    for j, (o, tr) in enumerate(zip(list(down_o) + [mid_o], list(down_t) + [mid_t])):
        assert o.shape == tr.shape
        _close(f"controlnet residual {j}", o, tr, 2e-4)


@pytest.mark.parametrize("nine", [False, True])
def test_controlnet_inpaint_loop_host_logic(emu, nine):
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    from oracle.controlnet_ref import build_ref_controlnet
    from oracle.loop_ref import denoise_controlnet_inpaint
    cfg = tiny_config("sd15")
    if nine:
        cfg.in_channels = 9
    ref = synth.build_ref_unet(cfg, rank=16)
    cn = build_ref_controlnet(tiny_config("sd15"), cond_block_out_channels=COND_CH)
    steps, B, h = 3, 2, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    lat = synth.synth_latents(B, h, h, seed=0)
    img, noise = synth.synth_latents(B, h, h, seed=7), synth.synth_latents(B, h, h, seed=8)
    ctrl = torch.rand(B, 3, 8 * h, 8 * h, generator=torch.Generator().manual_seed(11))
    mask = torch.zeros(B, 1, h, h)
    mask[:, :, h // 4: 3 * h // 4, h // 4: 3 * h // 4] = 1
    mil = img * (1 - mask) if nine else None
    kw = dict(guidance_scale=5.0, start_merge_step=1, conditioning_scale=0.5)
    want = denoise_controlnet_inpaint(ref, cn, make_scheduler("ddim"), lat, null, aug, txt, ctrl, img, noise, mask, steps, masked_image_latents=mil, **kw)
    den = B200Denoiser(_engine(ref), B200Scheduler("ddim"), use_cuda_graph=False)
    got = den.controlnet_inpaint(_cn_engine(cn), lat, null, aug, txt, ctrl, img, noise, mask, num_inference_steps=steps, masked_image_latents=mil, **kw)
    _close(f"controlnet+inpaint loop nine={nine}", got, want, 5e-4)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_processors_host_logic_against_reference_golden(emu, idx):
    """The drop-in processors (LoRA folding + caching, K/V packing, diffusers AttnProcessor protocol) on the golden vectors of the verbatim
    reference, with emulated kernels: 16-bit weights/activations (the processors insist on them), fp32 math."""
    import os
    from consistentid_b200.processors import Consistent_AttProcessor, Consistent_IPAttProcessor
    from oracle.unet_ref import Attention
    case = torch.load(os.path.join(os.path.dirname(__file__), "golden", "processors_golden.pt"))[idx]
    m, dt = case["meta"], torch.bfloat16
    a1 = Attention(m["C"], None, m["heads"], m["C"] // m["heads"]); a2 = Attention(m["C"], m["cad"], m["heads"], m["C"] // m["heads"])
    a1.load_state_dict(case["attn1"]); a2.load_state_dict(case["attn2"])
    p1 = Consistent_AttProcessor(hidden_size=m["C"], cross_attention_dim=None, rank=m["rank"])
    p2 = Consistent_IPAttProcessor(hidden_size=m["C"], cross_attention_dim=m["cad"], rank=m["rank"], scale=m["scale"], num_tokens=4)
    p1.load_state_dict(case["proc1"], strict=True); p2.load_state_dict(case["proc2"], strict=True)
    for mod in (a1, a2, p1, p2):
        mod.to(dt)
    a1.set_processor(p1); a2.set_processor(p2)
    x, ehs = case["x"].to(dt), case["ehs"].to(dt)
    for got, want in ((a1(x), case["y_self"]), (a2(x, encoder_hidden_states=ehs), case["y_cross"]), (a2(x, encoder_hidden_states=ehs), case["y_cross"])):
        _close("processor", got, want, 6e-2)
    p2.scale = 0.0
    assert (a2(x, encoder_hidden_states=ehs).float() - case["y_cross"]).abs().max().item() > 1e-3


@pytest.mark.parametrize("nine", [False, True])
def test_plain_inpaint_loop_host_logic(emu, nine):
    """pipelines/StableDIffusionInpaint_ConsistentID.py:305-359 (no ControlNet): B200Denoiser.inpaint vs oracle.loop_ref.denoise_inpaint."""
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    from oracle.loop_ref import denoise_inpaint
    cfg = tiny_config("sd15")
    if nine:
        cfg.in_channels = 9
    ref = synth.build_ref_unet(cfg, rank=16)
    steps, B, h = 3, 2, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    lat = synth.synth_latents(B, h, h, seed=0)
    img, noise = synth.synth_latents(B, h, h, seed=7), synth.synth_latents(B, h, h, seed=8)
    mask = torch.zeros(B, 1, h, h)
    mask[:, :, h // 4: 3 * h // 4, h // 4: 3 * h // 4] = 1
    mil = img * (1 - mask) if nine else None
    kw = dict(guidance_scale=5.0, start_merge_step=1)
    want = denoise_inpaint(ref, make_scheduler("ddim"), lat, null, aug, txt, img, noise, mask, steps, masked_image_latents=mil, **kw)
    den = B200Denoiser(_engine(ref), B200Scheduler("ddim"), use_cuda_graph=False)
    got = den.inpaint(lat, null, aug, txt, img, noise, mask, num_inference_steps=steps, masked_image_latents=mil, **kw)
    _close(f"plain inpaint loop nine={nine}", got, want, 5e-4)


def test_clip_vision_encoder_host_logic(emu):
    """consistentid_b200/clip.py on emulated kernels vs oracle/clip_ref.py (pinned on transformers): weight packing, the 264-row
    padded token buffers with masked pad keys, layer count (hidden_states[-2] skips the last layer)."""
    from consistentid_b200.clip import B200CLIPVisionEncoder
    from oracle import clip_ref
    from tests.test_clip_gpu import _weights
    C, heads, layers, inter, image = 128, 2, 3, 256, 42
    sd = _weights(C, heads, layers, inter, image, 14)
    x = torch.randn(2, 3, image, image, generator=torch.Generator().manual_seed(5))
    want = clip_ref.penultimate_hidden_state(sd, x, heads)
    enc = B200CLIPVisionEncoder(sd, num_attention_heads=heads, dtype=torch.float32, device="cpu")
    got = enc(x)
    assert got.shape == want.shape
    _close("clip vision encoder", got, want, 2e-4)


def test_prompt_cache_never_aliases_a_recycled_allocation(emu):
    """the drop-in ``unet(...)`` path keys its K/V cache on the prompt TENSOR.  A reference-style loop builds a fresh prompt
    tensor every step; once the old one is freed the allocator may hand the same address to a DIFFERENT prompt.  The cache entry must hold the
    keyed tensor (so that cannot happen) and must miss for a new tensor object even if it had the same address, shape and version."""
    cfg = tiny_config("sd15")
    ref = synth.build_ref_unet(cfg, rank=16)
    B, h = 1, cfg.sample_size
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(2 * B, h, h, seed=3)
    t = torch.tensor(601)
    eng = _engine(ref)
    with torch.no_grad():
        want_txt = ref(x, t, torch.cat([null, txt])).sample
        want_aug = ref(x, t, torch.cat([null, aug])).sample
    # reference-style loop: fresh torch.cat per step, alternating contents, previous tensors dropped
    for step in range(4):
        ehs = torch.cat([null, txt if step % 2 == 0 else aug])
        got = eng(x, t, ehs, cross_attention_kwargs={}).sample
        _close(f"step {step}", got, want_txt if step % 2 == 0 else want_aug, 2e-4)
        del ehs
    # the same tensor OBJECT twice hits the cache; an in-place edit (version bump) misses it
    ehs = torch.cat([null, txt])
    eng(x, t, ehs, cross_attention_kwargs={})
    key = eng._active_key
    eng(x, t, ehs, cross_attention_kwargs={})
    assert eng._active_key == key
    ehs.copy_(torch.cat([null, aug]))
    got = eng(x, t, ehs, cross_attention_kwargs={}).sample
    _close("after in-place edit", got, want_aug, 2e-4)
    # the cache entries hold the keyed tensors alive
    from consistentid_b200.weights import TensorIdent
    assert all(isinstance(i, TensorIdent) for ids in eng._ident.values() for i in ids if i is not None)
