"""CPU oracle: restatement of diffusers==0.23.0 ``ControlNetModel`` (SD1.5 ``control_v11p_sd15_inpaint`` topology) as driven by
pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:405-412.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED
(third-party dependency).

The ControlNet keeps diffusers' DEFAULT attention processors: plain attention over all 81 encoder tokens, no LoRA, no
id branch (component 12)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_ref import Attention, DownBlock, MidBlock, TimestepEmbedding, UNetConfig, get_timestep_embedding


class DefaultAttnProcessorRef:
    """diffusers ``AttnProcessor2_0``: q/k/v linears, SDPA, out linear."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        x = hidden_states
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b = x.shape[0]
        h = attn.heads
        q, k, v = attn.to_q(x), attn.to_k(ctx), attn.to_v(ctx)
        d = q.shape[-1] // h
        hf = lambda t: t.view(b, -1, h, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(hf(q), hf(k), hf(v), dropout_p=0.0)
        o = o.transpose(1, 2).reshape(b, -1, h * d).to(q.dtype)
        return attn.to_out[1](attn.to_out[0](o))


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, out_channels, cond_channels=3, block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(block_out_channels[-1], out_channels, 3, padding=1)   # zero-initialised in diffusers

    def forward(self, c):
        x = F.silu(self.conv_in(c))
        for blk in self.blocks:
            x = F.silu(blk(x))
        return self.conv_out(x)


class ControlNetRef(nn.Module):
    def __init__(self, cfg: UNetConfig, cond_block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(boc[0], 3, cond_block_out_channels)
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(boc[0], boc[0], 1)])
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(cfg, in_ch, out_ch, t.startswith("CrossAttn"), cfg.transformer_layers_per_block[i],
                                              cfg.num_attention_heads[i], add_downsample=not last))
            for _ in range(cfg.layers_per_block + (0 if last else 1)):
                self.controlnet_down_blocks.append(nn.Conv2d(out_ch, out_ch, 1))
        self.mid_block = MidBlock(cfg, boc[-1], cfg.transformer_layers_per_block[-1], cfg.num_attention_heads[-1])
        self.controlnet_mid_block = nn.Conv2d(boc[-1], boc[-1], 1)
        for m in self.modules():
            if isinstance(m, Attention):
                m.set_processor(DefaultAttnProcessorRef())

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0, return_dict=False):
        cfg = self.config
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=sample.device)
        elif timestep.ndim == 0:
            timestep = timestep[None].to(sample.device)
        t_emb = get_timestep_embedding(timestep.expand(sample.shape[0]), cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(sample.dtype))
        x = self.conv_in(sample) + self.controlnet_cond_embedding(controlnet_cond)
        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states, None)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states, None)
        down = tuple(conv(s) * conditioning_scale for s, conv in zip(skips, self.controlnet_down_blocks))
        mid = self.controlnet_mid_block(x) * conditioning_scale
        return down, mid


def build_ref_controlnet(cfg: UNetConfig, seed=4321, dtype=torch.float32, cond_block_out_channels=(16, 32, 96, 256)):
    """Synthetic weights; the zero-initialised output convs get small random weights so the residual path is exercised."""
    torch.manual_seed(seed)
    cn = ControlNetRef(cfg, cond_block_out_channels).to(dtype).eval()
    for p in cn.parameters():
        p.requires_grad_(False)
    return cn
