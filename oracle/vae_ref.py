"""CPU oracle: the VAE decode that follows the denoising loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: ``AutoencoderKL`` lives in the third-party
``diffusers==0.23.0`` (``requirements.txt:36``), absent from /root/reference and not installable here; the decoder is restated from
the published architecture of that release (``models/autoencoder_kl.py``, ``models/vae.py`` Decoder, ``UNetMidBlock2D``,
``UpDecoderBlock2D``, ``Attention`` with ``_from_deprecated_attn_block``).  Module / parameter names follow diffusers, so the ``decoder.*``
and ``post_quant_conv.*`` entries of a real ``vae`` state_dict load with ``strict=True``; the restated decoder has 49,490,179 parameters and
post_quant_conv 20, the published sizes of the SD VAE decoder.

Reference call sites:
  pipline_StableDiffusion_ConsistentID.py:586          image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]
  pipline_StableDiffusionXL_ConsistentID.py:669-684    same, after the fp16 -> fp32 upcast of the VAE (force_upcast)
Config (SD1.5 and SDXL VAEs): latent_channels 4, block_out_channels (128, 256, 512, 512), layers_per_block 2, norm_num_groups 32,
act_fn silu, scaling_factor 0.18215 (SD1.5) / 0.13025 (SDXL).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    name: str = "sd15_vae"


def sd15_vae_config():
    return VAEConfig()


def sdxl_vae_config():
    return VAEConfig(scaling_factor=0.13025, name="sdxl_vae")


def tiny_vae_config():
    return VAEConfig(block_out_channels=(64, 64, 128, 128), name="tiny_vae")


class _Resnet(nn.Module):
    """ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _AttnBlock(nn.Module):
    """diffusers Attention(channels, heads=1, dim_head=channels, bias=True, norm_num_groups=32, eps=1e-6, residual_connection=True,
    rescale_output_factor=1) with the default processor on a 4-D input: GroupNorm -> q,k,v -> softmax(q k^T / sqrt(C)) v -> to_out -> + input."""

    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.reshape(b, c, h * w)).transpose(1, 2)           # [B, HW, C]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax((q @ k.transpose(1, 2)).float() * (c ** -0.5), dim=-1).to(q.dtype)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class _MidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(ch, ch, groups), _Resnet(ch, ch, groups)])
        self.attentions = nn.ModuleList([_AttnBlock(ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Upsampler(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(x.dtype))


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, n_layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(n_layers)])
        self.upsamplers = nn.ModuleList([_Upsampler(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class DecoderRef(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = _MidBlock(boc[-1], g)
        rev = list(reversed(boc))
        blocks, prev = [], rev[0]
        for i, ch in enumerate(rev):
            blocks.append(_UpBlock(prev, ch, cfg.layers_per_block + 1, g, add_upsample=i != len(rev) - 1))
            prev = ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VAEDecodeRef(nn.Module):
    """``AutoencoderKL.decode`` without tiling/slicing: post_quant_conv (1x1) then the decoder.  ``decode_latents`` adds the pipelines'
    division by ``scaling_factor``."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.config = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = DecoderRef(cfg)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def decode_latents(self, latents):
        return self.decode(latents / self.config.scaling_factor)


def build_ref_vae(cfg: VAEConfig, seed=4321, dtype=torch.float32):
    torch.manual_seed(seed)
    m = VAEDecodeRef(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for n, p in m.named_parameters():        # non-trivial norms / biases so every term is exercised
        if p.ndim == 1:
            p.data = (1.0 if "norm" in n and n.endswith("weight") else 0.0) + 0.05 * torch.randn(p.shape, generator=g)
    return m.to(dtype)
