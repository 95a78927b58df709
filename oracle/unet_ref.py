"""CPU oracle: PyTorch restatement of ``diffusers==0.23.0`` UNet2DConditionModel semantics.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED for this file:
diffusers is a third-party dependency of the reference (``requirements.txt:36``),
absent from /root/reference and not installable here; the algorithm is restated
from SURVEY.md Appendix A.  Module / parameter names follow diffusers so a real
``unet`` state_dict would load with ``strict=True``.

Reference call sites this file stands in for:
  pipline_StableDiffusion_ConsistentID.py:552-557   (SD1.5 ``self.unet(...)``)
  pipline_StableDiffusionXL_ConsistentID.py:634-641 (SDXL, ``added_cond_kwargs``)
  pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:418-425 (residual inputs)
  attention.py:120-172,218-292 (fields of the ``attn`` module the processors touch)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- configs (A.1)
@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    num_attention_heads: Tuple[int, ...] = (8, 8, 8, 8)  # diffusers "attention_head_dim" = heads
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: Optional[int] = None
    projection_class_embeddings_input_dim: Optional[int] = None
    sample_size: int = 64
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    name: str = "sd15"

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4


def sd15_config(in_channels=4) -> UNetConfig:
    return UNetConfig(in_channels=in_channels, name="sd15")


def sdxl_config() -> UNetConfig:
    return UNetConfig(
        block_out_channels=(320, 640, 1280),
        down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
        transformer_layers_per_block=(1, 2, 10),
        num_attention_heads=(5, 10, 20),
        cross_attention_dim=2048,
        use_linear_projection=True,
        addition_embed_type="text_time",
        addition_time_embed_dim=256,
        projection_class_embeddings_input_dim=2816,
        sample_size=128,
        name="sdxl",
    )


def tiny_config(kind="sd15") -> UNetConfig:
    """Reduced-width configs with the same topology (CPU-CI sized).  Channels stay
    multiples of 64 so the same kernels/tilings are exercised."""
    if kind == "sd15":
        return UNetConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(2, 2, 4, 4),
                          cross_attention_dim=128, sample_size=32, name="tiny_sd15")
    c = sdxl_config()
    c.block_out_channels = (64, 128, 256)
    c.transformer_layers_per_block = (1, 1, 2)
    c.num_attention_heads = (1, 2, 4)
    c.cross_attention_dim = 128
    c.addition_time_embed_dim = 32
    c.projection_class_embeddings_input_dim = 64 + 6 * 32
    c.sample_size = 32
    c.name = "tiny_sdxl"
    return c


# --------------------------------------------------------------------------- embeddings (A.2)
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=True, downscale_freq_shift=0.0,
                           scale=1.0, max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


# --------------------------------------------------------------------------- attention (A.4)
class Attention(nn.Module):
    """Fields / helpers of diffusers ``Attention`` that the reference processors touch
    (attention.py:120-172, 218-292)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_softmax = False
        self.upcast_attention = False
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        raise NotImplementedError("the ConsistentID pipelines never pass an attention mask")

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        empty = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
        scores = torch.baddbmm(empty, query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        probs = scores.softmax(dim=-1)
        return probs.to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        x = self.attn1(self.norm1(x), encoder_hidden_states=None, **kw) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states, **kw) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, norm_num_groups,
                 use_linear_projection):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, encoder_hidden_states, cross_attention_kwargs=None):
        b, c, h, w = x.shape
        residual = x
        x = self.norm(x)
        if not self.use_linear_projection:
            x = self.proj_in(x)
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        else:
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
            x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states, cross_attention_kwargs)
        if not self.use_linear_projection:
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
            x = self.proj_out(x)
        else:
            x = self.proj_out(x)
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        return x + residual


# --------------------------------------------------------------------------- resnet / samplers (A.3)
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.output_scale_factor = 1.0

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = h + t
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        dtype = x.dtype
        if dtype == torch.bfloat16:  # diffusers 0.23 upcasts bf16 around interpolate (no numeric effect for nearest)
            x = x.to(torch.float32)
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        if dtype == torch.bfloat16:
            x = x.to(dtype)
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, in_ch, out_ch, has_attn, n_tf, heads, add_downsample):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, cfg.time_embed_dim,
                                              cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, out_ch // heads, out_ch, n_tf, cfg.cross_attention_dim,
                                                          cfg.norm_num_groups, cfg.use_linear_projection))
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, x, temb, ehs, kw):
        outs = ()
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_cross_attention:
                x = self.attentions[i](x, ehs, kw)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, ch, n_tf, heads):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, ch, n_tf, cfg.cross_attention_dim,
                                                            cfg.norm_num_groups, cfg.use_linear_projection)])
        self.resnets.append(ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))

    def forward(self, x, temb, ehs, kw):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs, kw)
        x = self.resnets[1](x, temb)
        return x


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, in_ch, out_ch, prev_out_ch, has_attn, n_tf, heads, add_upsample):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        n = cfg.layers_per_block + 1
        for i in range(n):
            skip = in_ch if i == n - 1 else out_ch
            rin = prev_out_ch if i == 0 else out_ch
            self.resnets.append(ResnetBlock2D(rin + skip, out_ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, out_ch // heads, out_ch, n_tf, cfg.cross_attention_dim,
                                                          cfg.norm_num_groups, cfg.use_linear_projection))
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x, skips, temb, ehs, kw):
        for i, res in enumerate(self.resnets):
            s = skips[-1]
            skips = skips[:-1]
            x = torch.cat([x, s], dim=1)
            x = res(x, temb)
            if self.has_cross_attention:
                x = self.attentions[i](x, ehs, kw)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionRef(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        self.in_channels = cfg.in_channels
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()   # registered BEFORE mid_block, as in diffusers: attn_processors order = down, up, mid
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownBlock(cfg, in_ch, out_ch, t.startswith("CrossAttn"),
                                              cfg.transformer_layers_per_block[i], cfg.num_attention_heads[i],
                                              add_downsample=(i != len(boc) - 1)))
        self.mid_block = MidBlock(cfg, boc[-1], cfg.transformer_layers_per_block[-1], cfg.num_attention_heads[-1])
        rev = list(reversed(boc))
        rev_heads = list(reversed(cfg.num_attention_heads))
        rev_tf = list(reversed(cfg.transformer_layers_per_block))
        out_ch = rev[0]
        for i, t in enumerate(cfg.up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cfg, in_ch, out_ch, prev, t.startswith("CrossAttn"), rev_tf[i], rev_heads[i],
                                          add_upsample=(i != len(boc) - 1)))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    # -- diffusers attn-processor plumbing (pipline_StableDiffusion_ConsistentID.py:152-174)
    @property
    def attn_processors(self):
        procs = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                procs[f"{name}.processor"] = m.processor
        return procs

    def set_attn_processor(self, processors):
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                m.set_processor(processors[f"{name}.processor"] if isinstance(processors, dict) else processors)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=True):
        cfg = self.config
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32 if isinstance(timestep, float) else torch.int64,
                                    device=sample.device)
        elif timestep.ndim == 0:
            timestep = timestep[None].to(sample.device)
        timesteps = timestep.expand(sample.shape[0])
        t_emb = get_timestep_embedding(timesteps, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        t_emb = t_emb.to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        if cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            time_embeds = get_timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim, cfg.flip_sin_to_cos,
                                                 cfg.freq_shift)
            time_embeds = time_embeds.reshape((text_embeds.shape[0], -1))
            add_embeds = torch.cat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add_embeds)
        x = self.conv_in(sample)
        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states, cross_attention_kwargs)
            skips += outs
        if down_block_additional_residuals is not None:
            skips = tuple(s + r for s, r in zip(skips, down_block_additional_residuals))
        x = self.mid_block(x, emb, encoder_hidden_states, cross_attention_kwargs)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for blk in self.up_blocks:
            n = len(blk.resnets)
            s, skips = skips[-n:], skips[:-n]
            x = blk(x, s, emb, encoder_hidden_states, cross_attention_kwargs)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)
