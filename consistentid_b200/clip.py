"""CLIP ViT image encoder on the GPU: ``image_encoder(pixels, output_hidden_states=True).hidden_states[-2]``
(pipline_StableDiffusion_ConsistentID.py:182-183, 202-203; ``laion/CLIP-ViT-H-14``: 1280 wide, 32 layers of which the last is never
needed, 16 heads of 80, 257 tokens; the pipeline pushes the face crop, the 5 facial-region crops and their zero images through it).
Takes the ``vision_model.*`` entries of a transformers ``CLIPVisionModelWithProjection`` state_dict.

Kernels reused from the hot path: ``cid_gemm`` (patch embedding as a GEMM over unfolded 14x14 patches with K 588 -> 640, fused-bias QKV
with the transposed-V epilogue, out-proj / fc2 with fused bias + residual, fc1 with the fused exact-erf GELU epilogue), ``cid_layernorm``,
``cid_attn_self_ragged`` (the 257 tokens live in 264-row buffers; the 7 pad keys are masked inside the flash kernel), ``cid_add_inplace``.
16-bit CUDA tensors only; there is no CPU / fp32 path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .lib import EPI_GELU, EPI_QKV


class B200CLIPVisionEncoder:
    def __init__(self, state_dict, num_attention_heads=16, hidden_act="gelu", dtype=torch.float16, device="cuda"):
        if hidden_act != "gelu":
            raise NotImplementedError("only hidden_act='gelu' (the laion ViT-H/14 the reference loads) is implemented")
        self.dtype, self.device, self.heads = dtype, torch.device(device), num_attention_heads
        ops.ensure_workspace(self.device)
        W = lambda n: state_dict[n].detach().to(device=self.device, dtype=dtype).contiguous()
        pw = W("vision_model.embeddings.patch_embedding.weight")                    # [C, 3, P, P]
        self.C, self.P = pw.shape[0], pw.shape[-1]
        if self.C % 64 or (self.C // num_attention_heads) % 8:
            raise ValueError(f"hidden size {self.C} / heads {num_attention_heads}: need C % 64 == 0 and head dim % 8 == 0")
        k = 3 * self.P * self.P
        self.Kp = (k + 63) // 64 * 64
        wp = torch.zeros((self.C, self.Kp), dtype=dtype, device=self.device)
        wp[:, :k] = pw.reshape(self.C, k)
        self.p = p = {"patch.w": wp, "cls": W("vision_model.embeddings.class_embedding"), "pos": W("vision_model.embeddings.position_embedding.weight")}
        self.n_tok = p["pos"].shape[0]
        self.n_pad = (self.n_tok + 7) // 8 * 8
        n_layers = 1 + max(int(n.split(".")[3]) for n in state_dict if n.startswith("vision_model.encoder.layers."))
        self.n_run = n_layers - 1                                                         # hidden_states[-2]: the last layer is skipped
        p["pre_layrnorm.g"], p["pre_layrnorm.b"] = W("vision_model.pre_layrnorm.weight"), W("vision_model.pre_layrnorm.bias")
        inter = None
        for i in range(self.n_run):
            b = f"vision_model.encoder.layers.{i}."
            for ln in ("layer_norm1", "layer_norm2"):
                p[f"{i}.{ln}.g"], p[f"{i}.{ln}.b"] = W(b + ln + ".weight"), W(b + ln + ".bias")
            p[f"{i}.qkv.w"] = torch.cat([W(b + f"self_attn.{q}_proj.weight") for q in "qkv"], 0).contiguous()
            p[f"{i}.qkv.b"] = torch.cat([W(b + f"self_attn.{q}_proj.bias") for q in "qkv"], 0).contiguous()
            p[f"{i}.o.w"], p[f"{i}.o.b"] = W(b + "self_attn.out_proj.weight"), W(b + "self_attn.out_proj.bias")
            p[f"{i}.fc1.w"], p[f"{i}.fc1.b"] = W(b + "mlp.fc1.weight"), W(b + "mlp.fc1.bias")
            p[f"{i}.fc2.w"], p[f"{i}.fc2.b"] = W(b + "mlp.fc2.weight"), W(b + "mlp.fc2.bias")
            inter = p[f"{i}.fc1.w"].shape[0]
        self.inter = inter
        if inter is not None and inter % 64:
            raise ValueError(f"intermediate size {inter} must be a multiple of 64")

    @torch.no_grad()
    def __call__(self, pixel_values):
        """pixel_values [B, 3, H, W] (H, W multiples of the patch size) -> hidden_states[-2] [B, 1 + n_patches, C]."""
        if not (pixel_values.is_cuda and pixel_values.dtype == self.dtype and pixel_values.ndim == 4):
            raise TypeError(f"B200CLIPVisionEncoder: expected a CUDA {self.dtype} tensor [B, 3, H, W] (no CPU/fp32 path)")
        p, C, H, Np = self.p, self.C, self.heads, self.n_pad
        d = C // H
        B = pixel_values.shape[0]
        new = lambda *s: torch.empty(s, dtype=self.dtype, device=self.device)
        cols = F.unfold(pixel_values, kernel_size=self.P, stride=self.P).transpose(1, 2)          # [B, n_patches, 3*P*P] (data movement only)
        n_patch = cols.shape[1]
        if n_patch + 1 != self.n_tok:
            raise ValueError(f"{n_patch} patches + class token != {self.n_tok} position embeddings")
        a = torch.zeros((B * n_patch, self.Kp), dtype=self.dtype, device=self.device)
        a[:, :cols.shape[2]] = cols.reshape(B * n_patch, -1)
        patches = ops.gemm(a, p["patch.w"], new(B * n_patch, C))
        x = torch.zeros((B, Np, C), dtype=self.dtype, device=self.device)
        x[:, 0] = p["cls"]
        x[:, 1:self.n_tok] = patches.view(B, n_patch, C)
        pos = torch.zeros((B, Np, C), dtype=self.dtype, device=self.device)
        pos[:, :self.n_tok] = p["pos"]
        ops.add_inplace(x, pos)
        M = B * Np
        x = ops.layernorm(x.view(M, C), p["pre_layrnorm.g"], p["pre_layrnorm.b"], new(M, C), M, C)
        h, qk, vt, ao, mid = new(M, C), new(M, 2 * C), new(B * H, d, Np), new(M, C), new(M, self.inter or C)
        for i in range(self.n_run):
            ops.layernorm(x, p[f"{i}.layer_norm1.g"], p[f"{i}.layer_norm1.b"], h, M, C)
            ops.gemm(h, p[f"{i}.qkv.w"], qk, bias=p[f"{i}.qkv.b"], epi=EPI_QKV, vt=vt, n_split=2 * C, heads=H, hdim=d, ntok=Np)
            ops.attn_self(qk[:, :C], qk[:, C:], vt, ao, B, H, Np, d, n_valid=self.n_tok)
            x = ops.gemm(ao, p[f"{i}.o.w"], new(M, C), bias=p[f"{i}.o.b"], residual=x)
            ops.layernorm(x, p[f"{i}.layer_norm2.g"], p[f"{i}.layer_norm2.b"], h, M, C)
            ops.gemm(h, p[f"{i}.fc1.w"], mid, bias=p[f"{i}.fc1.b"], epi=EPI_GELU)
            x = ops.gemm(mid, p[f"{i}.fc2.w"], new(M, C), bias=p[f"{i}.fc2.b"], residual=x)
        return x.view(B, Np, C)[:, :self.n_tok].contiguous()
