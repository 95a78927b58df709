"""Embedding producers on the GPU: the modules that write the hot path's ``encoder_hidden_states``.

Drop-in counterparts of the reference classes - same names, constructor arguments, call signatures and ``state_dict`` keys, so the
``image_proj`` / ``FacialEncoder`` sections of a ConsistentID checkpoint (checkpoint.py) load with ``load_state_dict(strict=True)``:

  ProjPlusModel   functions.py:494-528   (id embedding [B,512] + CLIP patch features [B,257,1280]) -> 4 id tokens [B,4,cad]
  FacialEncoder   attention.py:72-88     AttentionMLP (functions.py:530-592) over 5 facial-region crops + FuseModule (attention.py:10-48)
                                          writing the fused rows at the <|facial|> token positions of the text embedding
  assemble_prompts pipline_StableDiffusion_ConsistentID.py:494-507  -> the (null, augmented, text_only) [1,81,cad] prompt tensors

All arithmetic runs in the library's kernels: the wide projections over the 257-row CLIP features (``proj_in``, ``to_kv``) on the
tcgen05 GEMM, LayerNorms through ``cid_layernorm_rows`` (written straight into the concatenated key/value input), the 1-4 latent rows
through ``cid_skinny_linear`` (bias / GELU / residual fused) and ``cid_perceiver_attn``.  torch is used for allocation, the boolean-mask
gather/scatter of FuseModule (data movement) and the optional ``x + scale*out`` shortcut on 4 rows.  16-bit CUDA tensors only; there is
no CPU path.
"""
from __future__ import annotations

import os

import torch

from . import ops

DIM_HEAD = 64


def _need_cuda16(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in (torch.float16, torch.bfloat16)):
        raise TypeError(f"{name}: expected a CUDA fp16/bf16 tensor (the B200 embedding producers have no CPU/fp32 path)")


class _Weights:
    """state_dict holder with the reference's strict-loading behaviour."""

    def __init__(self, shapes, dtype, device):
        self._shapes, self.dtype, self.device = shapes, dtype, torch.device(device)
        self._w = {}

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._shapes if k not in sd]
        unexpected = [k for k in sd if k not in self._shapes]
        bad = [k for k in self._shapes if k in sd and tuple(sd[k].shape) != tuple(self._shapes[k])]
        if bad or (strict and (missing or unexpected)):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:6]} ({len(missing)}), unexpected {unexpected[:6]} "
                               f"({len(unexpected)}), size mismatch {bad[:6]} ({len(bad)})")
        for k in self._shapes:
            if k in sd:
                self._w[k] = sd[k].detach().to(device=self.device, dtype=self.dtype).contiguous()
        return self

    def state_dict(self):
        return dict(self._w)

    def __getitem__(self, k):
        try:
            return self._w[k]
        except KeyError:
            raise RuntimeError(f"parameter '{k}' has not been loaded (call load_state_dict first)") from None


def _perceiver_shapes(p, dim, depth, heads, emb, out_dim, ff_mult=4):
    inner = heads * DIM_HEAD
    s = {f"{p}proj_in.weight": (dim, emb), f"{p}proj_in.bias": (dim,), f"{p}proj_out.weight": (out_dim, dim), f"{p}proj_out.bias": (out_dim,),
         f"{p}norm_out.weight": (out_dim,), f"{p}norm_out.bias": (out_dim,)}
    for i in range(depth):
        a, f = f"{p}layers.{i}.0.", f"{p}layers.{i}.1."
        s.update({a + "norm1.weight": (dim,), a + "norm1.bias": (dim,), a + "norm2.weight": (dim,), a + "norm2.bias": (dim,),
                  a + "to_q.weight": (inner, dim), a + "to_kv.weight": (2 * inner, dim), a + "to_out.weight": (dim, inner),
                  f + "0.weight": (dim,), f + "0.bias": (dim,), f + "1.weight": (dim * ff_mult, dim), f + "3.weight": (dim, dim * ff_mult)})
    return s


def _perceiver_stack(w, p, depth, heads, latents, x):
    """FacePerceiverResampler.forward / AttentionMLP.forward body (functions.py:486-492, 583-592).
    latents [B, L, dim] (consumed), x [B, n, emb] -> [B, L, out_dim]."""
    B, n, emb = x.shape
    L, dim = latents.shape[1], latents.shape[2]
    inner = heads * DIM_HEAD
    dt, dev = x.dtype, x.device
    new = lambda *s: torch.empty(s, dtype=dt, device=dev)
    xp = ops.gemm(x.reshape(B * n, emb), w[p + "proj_in.weight"], new(B * n, dim), bias=w[p + "proj_in.bias"])
    lat = latents.reshape(B * L, dim).contiguous()
    kv_in, lat_n, q, kv, o = new(B * (n + L), dim), new(B * L, dim), new(B * L, inner), new(B * (n + L), 2 * inner), new(B * L, inner)
    h0, h1 = new(B * L, dim), new(B * L, 4 * dim)
    for i in range(depth):
        a, f = f"{p}layers.{i}.0.", f"{p}layers.{i}.1."
        # kv_input = cat(norm1(x), norm2(latents)) built in place
        ops.layernorm_rows(xp, w[a + "norm1.weight"], w[a + "norm1.bias"], kv_in, B * n, dim, rows_per_group=n, y_group_rows=n + L, y_row0=0)
        ops.layernorm_rows(lat, w[a + "norm2.weight"], w[a + "norm2.bias"], kv_in, B * L, dim, rows_per_group=L, y_group_rows=n + L, y_row0=n)
        ops.layernorm_rows(lat, w[a + "norm2.weight"], w[a + "norm2.bias"], lat_n, B * L, dim)
        ops.skinny_linear(lat_n, w[a + "to_q.weight"], None, q, B * L, inner, dim)
        ops.gemm(kv_in, w[a + "to_kv.weight"], kv)
        ops.perceiver_attn(q, kv, o, B, L, n + L, heads)
        ops.skinny_linear(o, w[a + "to_out.weight"], None, lat, B * L, dim, inner, accumulate=True)             # + latents
        ops.layernorm_rows(lat, w[f + "0.weight"], w[f + "0.bias"], h0, B * L, dim)
        ops.skinny_linear(h0, w[f + "1.weight"], None, h1, B * L, 4 * dim, dim)
        ops.skinny_linear(h1, w[f + "3.weight"], None, lat, B * L, dim, 4 * dim, act_in="gelu", accumulate=True)  # + latents
    out_dim = w[p + "proj_out.weight"].shape[0]
    po = ops.skinny_linear(lat, w[p + "proj_out.weight"], w[p + "proj_out.bias"], new(B * L, out_dim), B * L, out_dim, dim)
    y = ops.layernorm_rows(po, w[p + "norm_out.weight"], w[p + "norm_out.bias"], new(B * L, out_dim), B * L, out_dim)
    return y.reshape(B, L, out_dim)


class _GraphedStack:
    """The perceiver stack is ~15 tiny launches per layer on 1-4 latent rows: launch-bound.  Per input shape it is captured once in a CUDA
    graph (static input/output buffers) and replayed; CID_EMBED_GRAPH=0 keeps the eager launch sequence."""

    def __init__(self, w, prefix, depth, heads):
        self.w, self.prefix, self.depth, self.heads = w, prefix, depth, heads
        self.cache = {}
        self.enabled = os.environ.get("CID_EMBED_GRAPH", "1") != "0"

    def __call__(self, latents, x):
        if not self.enabled:
            return _perceiver_stack(self.w, self.prefix, self.depth, self.heads, latents.clone(), x)
        key = (tuple(latents.shape), tuple(x.shape), x.dtype)
        ent = self.cache.get(key)
        if ent is None:
            lat_s, x_s = latents.clone().contiguous(), x.clone().contiguous()
            _perceiver_stack(self.w, self.prefix, self.depth, self.heads, lat_s.clone(), x_s)      # warm-up outside the capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out_s = _perceiver_stack(self.w, self.prefix, self.depth, self.heads, lat_s, x_s)   # consumes lat_s in place
            ent = self.cache[key] = (graph, lat_s, x_s, out_s)
        graph, lat_s, x_s, out_s = ent
        lat_s.copy_(latents); x_s.copy_(x)
        graph.replay()
        return out_s.clone()


class ProjPlusModel:
    """functions.py:494-528 (``image_proj_model`` of the pipelines, pipline_StableDiffusion_ConsistentID.py:89-94)."""

    def __init__(self, cross_attention_dim=768, id_embeddings_dim=512, clip_embeddings_dim=1280, num_tokens=4, dtype=torch.float16, device="cuda"):
        self.cross_attention_dim, self.num_tokens = cross_attention_dim, num_tokens
        self.heads, self.depth = cross_attention_dim // 64, 4
        d, c = id_embeddings_dim, cross_attention_dim
        shapes = {"proj.0.weight": (2 * d, d), "proj.0.bias": (2 * d,), "proj.2.weight": (c * num_tokens, 2 * d), "proj.2.bias": (c * num_tokens,),
                  "norm.weight": (c,), "norm.bias": (c,)}
        shapes.update(_perceiver_shapes("perceiver_resampler.", c, self.depth, self.heads, clip_embeddings_dim, c))
        self.w = _Weights(shapes, dtype, device)
        ops.ensure_workspace(self.w.device)
        self._stack = _GraphedStack(self.w, "perceiver_resampler.", self.depth, self.heads)

    def load_state_dict(self, sd, strict=True):
        self.w.load_state_dict(sd, strict)
        self._stack.cache.clear()            # captured graphs point at the previous weight tensors
        return self

    def state_dict(self):
        return self.w.state_dict()

    @torch.no_grad()
    def __call__(self, id_embeds, clip_embeds, shortcut=False, scale=1.0):
        _need_cuda16(id_embeds, "id_embeds"); _need_cuda16(clip_embeds, "clip_embeds")
        w, c, T = self.w, self.cross_attention_dim, self.num_tokens
        idv = id_embeds.reshape(-1, id_embeds.shape[-1]).contiguous()
        B, d = idv.shape
        new = lambda *s: torch.empty(s, dtype=idv.dtype, device=idv.device)
        h = ops.skinny_linear(idv, w["proj.0.weight"], w["proj.0.bias"], new(B, 2 * d), B, 2 * d, d)
        t = ops.skinny_linear(h, w["proj.2.weight"], w["proj.2.bias"], new(B, c * T), B, c * T, 2 * d, act_in="gelu")
        x = ops.layernorm_rows(t.view(B * T, c), w["norm.weight"], w["norm.bias"], new(B * T, c), B * T, c).view(B, T, c)
        out = self._stack(x, clip_embeds.contiguous())
        return torch.add(x, out, alpha=float(scale)) if shortcut else out


class FacialEncoder:
    """attention.py:72-88: ``visual_projection`` = AttentionMLP(dim 1024, depth 8, 16 heads, one latent) and ``fuse_module`` = FuseModule."""

    def __init__(self, image_CLIPModel_encoder=None, embedding_dim=1280, output_dim=768, embed_dim=768, dtype=torch.float16, device="cuda",
                 dim=1024, depth=8, heads=16):
        self.dim, self.depth, self.heads, self.embed_dim = dim, depth, heads, embed_dim
        D = embed_dim
        shapes = {"visual_projection.latents": (1, 1, dim)}
        shapes.update(_perceiver_shapes("visual_projection.", dim, depth, heads, embedding_dim, output_dim))
        for m, cin in (("mlp1", 2 * D), ("mlp2", D)):
            p = f"fuse_module.{m}."
            shapes.update({p + "layernorm.weight": (cin,), p + "layernorm.bias": (cin,), p + "fc1.weight": (D, cin), p + "fc1.bias": (D,),
                           p + "fc2.weight": (D, D), p + "fc2.bias": (D,)})
        shapes.update({"fuse_module.layer_norm.weight": (D,), "fuse_module.layer_norm.bias": (D,)})
        self.w = _Weights(shapes, dtype, device)
        ops.ensure_workspace(self.w.device)
        self._stack = _GraphedStack(self.w, "visual_projection.", depth, heads)

    def load_state_dict(self, sd, strict=True):
        self.w.load_state_dict(sd, strict)
        self._stack.cache.clear()            # captured graphs point at the previous weight tensors
        return self

    def state_dict(self):
        return self.w.state_dict()

    @torch.no_grad()
    def visual_projection(self, x):
        """AttentionMLP.forward: [n, tokens, embedding_dim] -> [n, 1, output_dim]."""
        _need_cuda16(x, "multi_image_embeds")
        lat = self.w["visual_projection.latents"].repeat(x.shape[0], 1, 1)
        return self._stack(lat, x.contiguous())

    def _mlp(self, p, x, residual):
        w, m, D = self.w, x.shape[0], self.embed_dim
        a = ops.layernorm_rows(x, w[p + "layernorm.weight"], w[p + "layernorm.bias"], torch.empty_like(x), m, x.shape[1])
        h = ops.skinny_linear(a, w[p + "fc1.weight"], w[p + "fc1.bias"], torch.empty((m, D), dtype=x.dtype, device=x.device), m, D, x.shape[1])
        out = residual.clone()
        return ops.skinny_linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"], out, m, D, D, act_in="gelu", accumulate=True)

    @torch.no_grad()
    def __call__(self, prompt_embeds, multi_image_embeds, class_tokens_mask, valid_id_mask):
        _need_cuda16(prompt_embeds, "prompt_embeds"); _need_cuda16(multi_image_embeds, "multi_image_embeds")
        bs, n, tl, idim = multi_image_embeds.shape
        D = prompt_embeds.shape[-1]
        id_embeds = self.visual_projection(multi_image_embeds.reshape(bs * n, tl, idim)).reshape(bs * n, -1).to(prompt_embeds.dtype)
        tok = class_tokens_mask.reshape(-1).to(prompt_embeds.device).nonzero().reshape(-1)
        val = valid_id_mask.reshape(-1).to(prompt_embeds.device).nonzero().reshape(-1)
        assert tok.numel() == val.numel(), f"{tok.numel()} != {val.numel()}"          # attention.py:44
        flat = prompt_embeds.reshape(-1, D).clone()
        if tok.numel():
            rows, ids = flat[tok], id_embeds[val]
            fused = self._mlp("fuse_module.mlp1.", torch.cat([rows, ids], dim=-1).contiguous(), rows)       # mlp1(cat) + prompt rows
            fused = self._mlp("fuse_module.mlp2.", fused, fused)                                             # residual MLP
            fused = ops.layernorm_rows(fused, self.w["fuse_module.layer_norm.weight"], self.w["fuse_module.layer_norm.bias"],
                                       torch.empty_like(fused), fused.shape[0], D)
            flat[tok] = fused
        return flat.reshape(prompt_embeds.shape)


def assemble_prompts(facial_text, uncond_facial_text, text_only, id_tokens, uncond_id_tokens):
    """pipline_StableDiffusion_ConsistentID.py:494-507: (null, augmented, text_only), each [1, 77+4, cad] - the three prompt tensors
    ``B200Denoiser.__call__`` takes."""
    return (torch.cat([uncond_facial_text, uncond_id_tokens], dim=1), torch.cat([facial_text, id_tokens], dim=1),
            torch.cat([text_only, id_tokens], dim=1))
