"""VAE decode oracle (oracle/vae_ref.py): structure anchored on the published size of the SD VAE decoder and on the
diffusers parameter names; host-side checks of the engine's state_dict validation need no GPU."""
import pytest
import torch

from oracle.vae_ref import VAEDecodeRef, build_ref_vae, sd15_vae_config, sdxl_vae_config, tiny_vae_config


def test_decoder_parameter_inventory():
    with torch.device("meta"):
        m = VAEDecodeRef(sd15_vae_config())
    assert sum(p.numel() for p in m.decoder.parameters()) == 49_490_179          # AutoencoderKL decoder of SD1.5 / SDXL
    assert sum(p.numel() for p in m.post_quant_conv.parameters()) == 20
    keys = set(m.state_dict())
    for k in ("post_quant_conv.weight", "decoder.conv_in.weight", "decoder.mid_block.attentions.0.group_norm.weight",
              "decoder.mid_block.attentions.0.to_q.weight", "decoder.mid_block.attentions.0.to_out.0.bias", "decoder.mid_block.resnets.1.conv2.weight",
              "decoder.up_blocks.0.upsamplers.0.conv.weight", "decoder.up_blocks.2.resnets.0.conv_shortcut.weight",
              "decoder.up_blocks.3.resnets.2.norm2.bias", "decoder.conv_norm_out.weight", "decoder.conv_out.bias"):
        assert k in keys, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys              # the last block does not upsample
    assert "decoder.up_blocks.0.resnets.0.conv_shortcut.weight" not in keys       # 512 -> 512 has no shortcut conv
    assert sdxl_vae_config().scaling_factor == 0.13025 and sd15_vae_config().scaling_factor == 0.18215


def test_decode_shapes_and_scaling():
    vae = build_ref_vae(tiny_vae_config())
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = vae.decode_latents(z)
        b = vae.decode(z / vae.config.scaling_factor)
    assert a.shape == (2, 3, 64, 64) and torch.equal(a, b)
    # images are independent (no cross-sample op): batch of 2 == two batch-1 decodes
    with torch.no_grad():
        one = vae.decode_latents(z[1:2])
    assert torch.allclose(a[1:2], one, atol=1e-5, rtol=1e-5)
