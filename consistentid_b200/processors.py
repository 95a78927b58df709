"""Drop-in replacements for the reference's attention processors behind the diffusers ``AttnProcessor`` protocol.

``Consistent_AttProcessor`` / ``Consistent_IPAttProcessor`` keep the reference's constructor signatures, parameter names
(``to_{q,k,v,out}_lora.{down,up}.weight``, ``to_k_ip.weight``, ``to_v_ip.weight`` -> the checkpoint's ``adapter_modules``
loads with ``strict=True``, pipline_StableDiffusion_ConsistentID.py:143-144), the mutable ``scale`` attribute and the call
signature ``processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, [scale], temb=None)``
(attention.py:92-117, 179-215), so ``unet.set_attn_processor({name: processor})`` (:174) works unchanged with a diffusers
UNet.  The arithmetic of ``__call__`` runs in libcidb200: LoRA folded into the base projections (cached, re-folded when a
weight changes), fused QKV GEMM with transposed-V epilogue, tcgen05 flash self-attention / decoupled dual-softmax
cross-attention, output projection with fused bias.  There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .lib import EPI_QKV
from .weights import TensorIdent, fold_lora, same_tensors


class LoRALinearLayer(nn.Module):
    """Parameter container with the layout of diffusers 0.23 ``LoRALinearLayer`` (down: [rank, in], up: [out, rank])."""

    def __init__(self, in_features, out_features, rank=4, network_alpha=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        self.network_alpha, self.rank = network_alpha, rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)


def _idents(*ts):
    return tuple(TensorIdent(t) for t in ts)


class _ProcessorBase(nn.Module):
    def _folded(self, tag, base_w, lora):
        """base_w + lora_scale * up @ down, cached until any of the three tensors is modified."""
        srcs = (base_w, lora.down.weight, lora.up.weight)
        hit = self._cache.get(tag)
        if hit is None or hit[1] != self.lora_scale or not same_tensors(hit[0], srcs):
            w = fold_lora(base_w.detach(), lora.down.weight.detach().to(base_w.dtype), lora.up.weight.detach().to(base_w.dtype),
                          self.lora_scale, lora.network_alpha).contiguous()
            self._cache[tag] = hit = (_idents(*srcs), self.lora_scale, w)
        return hit[2]

    @staticmethod
    def _prep(attn, hidden_states):
        if attn.spatial_norm is not None or attn.group_norm is not None or attn.norm_cross:
            raise NotImplementedError("spatial_norm / group_norm / norm_cross are unused by the SD / SDXL UNets")
        if not hidden_states.is_cuda or hidden_states.dtype not in (torch.float16, torch.bfloat16):
            raise RuntimeError("consistentid_b200 processors run on CUDA fp16/bf16 tensors only (no CPU / fp32 fallback)")
        shape4 = None
        x = hidden_states
        if x.ndim == 4:
            b, c, h, w = x.shape
            shape4 = (b, c, h, w)
            x = x.view(b, c, h * w).transpose(1, 2)
        return x.contiguous(), shape4

    @staticmethod
    def _finish(attn, y, residual, shape4):
        if shape4 is not None:
            b, c, h, w = shape4
            y = y.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            y = y + residual
        if attn.rescale_output_factor != 1.0:
            y = y / attn.rescale_output_factor
        return y


class Consistent_AttProcessor(_ProcessorBase):
    """LoRA'd self-attention (attention.py:90-174)."""

    def __init__(self, hidden_size=None, cross_attention_dim=None, rank=4, network_alpha=None, lora_scale=1.0):
        super().__init__()
        self.rank, self.lora_scale = rank, lora_scale
        kv = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_lora = LoRALinearLayer(kv, hidden_size, rank, network_alpha)
        self.to_v_lora = LoRALinearLayer(kv, hidden_size, rank, network_alpha)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self._cache = {}

    @torch.no_grad()
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if encoder_hidden_states is not None or attention_mask is not None:
            raise NotImplementedError("Consistent_AttProcessor is installed on attn1 (self-attention, no mask) only")
        residual = hidden_states
        x, shape4 = self._prep(attn, hidden_states)
        B, N, C = x.shape
        H = attn.heads
        d = C // H
        M = B * N
        wq = self._folded("q", attn.to_q.weight, self.to_q_lora)
        wk = self._folded("k", attn.to_k.weight, self.to_k_lora)
        wv = self._folded("v", attn.to_v.weight, self.to_v_lora)
        wkey = self._cache.get("qkv_key")                 # the folded weights are this object's own cache entries: identity is enough
        if wkey is None or wkey[0] is not wq or wkey[1] is not wk or wkey[2] is not wv:
            self._cache["qkv"] = torch.cat([wq, wk, wv], 0).contiguous()
            self._cache["qkv_key"] = (wq, wk, wv)
        wo = self._folded("o", attn.to_out[0].weight, self.to_out_lora)
        x2 = x.view(M, C)
        qk = torch.empty((M, 2 * C), dtype=x.dtype, device=x.device)
        vt = torch.empty((B * H, d, N), dtype=x.dtype, device=x.device)
        ops.gemm(x2, self._cache["qkv"], qk, epi=EPI_QKV, vt=vt, n_split=2 * C, heads=H, hdim=d, ntok=N)
        ao = torch.empty((M, C), dtype=x.dtype, device=x.device)
        ops.attn_self(qk[:, :C], qk[:, C:], vt, ao, B, H, N, d)
        y = torch.empty((M, C), dtype=x.dtype, device=x.device)
        ops.gemm(ao, wo, y, bias=attn.to_out[0].bias)
        return self._finish(attn, y.view(B, N, C), residual, shape4)


class Consistent_IPAttProcessor(_ProcessorBase):
    """LoRA'd text cross-attention + decoupled id-token branch (attention.py:177-294)."""

    def __init__(self, hidden_size, cross_attention_dim=None, rank=4, network_alpha=None, lora_scale=1.0, scale=1.0, num_tokens=4):
        super().__init__()
        self.rank, self.lora_scale, self.num_tokens = rank, lora_scale, num_tokens
        self.hidden_size, self.cross_attention_dim, self.scale = hidden_size, cross_attention_dim, scale
        kv = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_lora = LoRALinearLayer(kv, hidden_size, rank, network_alpha)
        self.to_v_lora = LoRALinearLayer(kv, hidden_size, rank, network_alpha)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_ip = nn.Linear(kv, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(kv, hidden_size, bias=False)
        self._cache = {}

    @torch.no_grad()
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, scale=1.0, temb=None):
        if encoder_hidden_states is None or attention_mask is not None:
            raise NotImplementedError("Consistent_IPAttProcessor is installed on attn2 (cross-attention, no mask) only")
        residual = hidden_states
        x, shape4 = self._prep(attn, hidden_states)
        B, N, C = x.shape
        H = attn.heads
        d = C // H
        M = B * N
        ehs = encoder_hidden_states
        n_ip = self.num_tokens
        n_text = ehs.shape[1] - n_ip
        cad = ehs.shape[2]
        wq = self._folded("q", attn.to_q.weight, self.to_q_lora)
        wk = self._folded("k", attn.to_k.weight, self.to_k_lora)
        wv = self._folded("v", attn.to_v.weight, self.to_v_lora)
        wo = self._folded("o", attn.to_out[0].weight, self.to_out_lora)
        # K/V depend only on the prompt: cached per (prompt tensor OBJECT, weights).  The entry holds the prompt tensor itself, so a
        # different prompt can never alias it through a recycled allocation; a loop that re-creates the prompt tensor every step (the
        # reference's torch.cat at pipline_StableDiffusion_ConsistentID.py:542-549) recomputes K/V every step (4 small GEMMs + a pack).
        for wt in (self.to_k_ip.weight, self.to_v_ip.weight):
            if wt.dtype != x.dtype:
                raise TypeError(f"to_k_ip / to_v_ip are {wt.dtype} but the activations are {x.dtype}: move the processor with .to(dtype) "
                                "as the reference does (pipline_StableDiffusion_ConsistentID.py:166-172)")
        kv_srcs = (ehs, self.to_k_ip.weight, self.to_v_ip.weight)
        kv_hit = self._cache.get("kv_key")
        if kv_hit is None or kv_hit[1] is not wk or kv_hit[2] is not wv or not same_tensors(kv_hit[0], kv_srcs):
            text = ehs[:, :n_text].reshape(B * n_text, cad).contiguous()
            ip = ehs[:, n_text:].reshape(B * n_ip, cad).contiguous()
            new = lambda r: torch.empty((r, C), dtype=x.dtype, device=x.device)
            kt, vt_, ki, vi = new(B * n_text), new(B * n_text), new(B * n_ip), new(B * n_ip)
            ops.gemm(text, wk, kt); ops.gemm(text, wv, vt_)
            ops.gemm(ip, self.to_k_ip.weight.detach().contiguous(), ki); ops.gemm(ip, self.to_v_ip.weight.detach().contiguous(), vi)
            k_cat = torch.empty((B, 96, C), dtype=x.dtype, device=x.device)
            vt_cat = torch.empty((B * H, d, 96), dtype=x.dtype, device=x.device)
            ops.pack_cross_kv(kt, vt_, ki, vi, k_cat, vt_cat, B, C, H, n_text, n_ip)
            self._cache["kv"], self._cache["kv_key"] = (k_cat, vt_cat), (_idents(*kv_srcs), wk, wv)
        k_cat, vt_cat = self._cache["kv"]
        q = torch.empty((M, C), dtype=x.dtype, device=x.device)
        ops.gemm(x.view(M, C), wq, q)
        ao = torch.empty((M, C), dtype=x.dtype, device=x.device)
        ops.attn_cross(q, k_cat, vt_cat, ao, B, H, N, d, n_text, n_ip, float(self.scale))
        y = torch.empty((M, C), dtype=x.dtype, device=x.device)
        ops.gemm(ao, wo, y, bias=attn.to_out[0].bias)
        return self._finish(attn, y.view(B, N, C), residual, shape4)
