"""GPU parity of the VAE decoder (consistentid_b200/vae.py) against oracle/vae_ref.py (fp32 CPU truth, 16-bit eager GPU error bar)."""
import pytest
import torch

from oracle.vae_ref import build_ref_vae, sd15_vae_config, tiny_vae_config
from tests.test_unet_gpu import _cmp


def _engine(ref, cfg, dtype):
    from consistentid_b200.vae import B200VAEDecoder
    return B200VAEDecoder(ref.state_dict(), scaling_factor=cfg.scaling_factor, block_out_channels=cfg.block_out_channels,
                          layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups, dtype=dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,B,h,dtype", [("tiny", 2, 8, torch.float16), ("tiny", 1, 16, torch.bfloat16), ("sd15", 1, 32, torch.float16)])
def test_vae_decode_parity(kind, B, h, dtype):
    cfg = tiny_vae_config() if kind == "tiny" else sd15_vae_config()
    ref = build_ref_vae(cfg)
    z = torch.randn(B, 4, h, h, generator=torch.Generator().manual_seed(11)) * cfg.scaling_factor * 3.0
    with torch.no_grad():
        truth = ref.decode_latents(z)
        ref16 = build_ref_vae(cfg, dtype=dtype).cuda()
        eager = ref16.decode_latents(z.cuda().to(dtype))
    eng = _engine(ref, cfg, dtype)
    out = eng.decode_latents(z.cuda().to(dtype))
    out2 = eng.decode((z / cfg.scaling_factor).cuda().to(dtype), return_dict=False)[0]
    torch.cuda.synchronize()
    assert out.shape == (B, 3, 8 * h, 8 * h)
    _cmp(f"vae decode {kind} B{B} {h}x{h} {dtype}", out, truth, eager)
    _cmp(f"vae decode (pre-scaled latents) {kind} {dtype}", out2, truth, eager)


@pytest.mark.gpu
def test_vae_rejects_cpu_and_fp32_inputs():
    cfg = tiny_vae_config()
    eng = _engine(build_ref_vae(cfg), cfg, torch.float16)
    with pytest.raises(TypeError):
        eng.decode(torch.randn(1, 4, 8, 8))
    with pytest.raises(TypeError):
        eng.decode(torch.randn(1, 4, 8, 8, device="cuda"))
