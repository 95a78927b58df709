"""CPU oracle: the embedding producers that write the hot path's ``encoder_hidden_states``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Functional restatement over plain state_dicts whose keys are the
reference modules' parameter names, so the ``image_proj`` / ``FacialEncoder`` sections of a ConsistentID checkpoint feed it
unchanged.  PINNED: tests/golden/embed_golden.pt holds outputs of the reference's own classes (functions.py / attention.py
imported verbatim by tests/golden/make_embed_golden.py); tests/test_embed_cpu.py checks this file against them.

Follows:
  functions.py:389-397   FeedForward      LN -> Linear(no bias) -> GELU -> Linear(no bias)
  functions.py:407-455   PerceiverAttention  keys/values over cat(LN1(x), LN2(latents)), queries from LN2(latents),
                         q and k each scaled by dim_head^-1/4, softmax in fp32
  functions.py:457-492   FacePerceiverResampler
  functions.py:494-528   ProjPlusModel    id embedding -> 4 tokens -> resampler over CLIP patch features (+ optional shortcut)
  functions.py:530-592   AttentionMLP     (apply_pos_emb=False, num_latents_mean_pooled=0: the configuration FacialEncoder builds)
  attention.py:10-48     FuseModule       masked gather of the <|facial|> rows, two MLPs, LayerNorm, masked scatter
  attention.py:50-70     MLP
  attention.py:72-88     FacialEncoder
  pipline_StableDiffusion_ConsistentID.py:494-507  assembly of [77 fused text rows | 4 id rows]
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DIM_HEAD = 64     # both perceiver stacks use dim_head=64 (functions.py:505-507, 534-535)


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def perceiver_attention(sd, p, x, latents):
    b, l, _ = latents.shape
    xn, ln = _ln(sd, p + ".norm1", x), _ln(sd, p + ".norm2", latents)
    q = F.linear(ln, sd[p + ".to_q.weight"])
    k, v = F.linear(torch.cat([xn, ln], dim=1), sd[p + ".to_kv.weight"]).chunk(2, dim=-1)
    heads = q.shape[-1] // DIM_HEAD
    split = lambda t: t.reshape(b, t.shape[1], heads, DIM_HEAD).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    s = DIM_HEAD ** -0.25
    w = (q * s) @ (k * s).transpose(-2, -1)
    w = torch.softmax(w.float(), dim=-1).to(w.dtype)
    o = (w @ v).transpose(1, 2).reshape(b, l, heads * DIM_HEAD)
    return F.linear(o, sd[p + ".to_out.weight"])


def feed_forward(sd, p, x):
    h = F.linear(_ln(sd, p + ".0", x), sd[p + ".1.weight"])
    return F.linear(F.gelu(h), sd[p + ".3.weight"])


def _perceiver_stack(sd, p, latents, x):
    depth = 1 + max(int(k[len(p) + len(".layers."):].split(".")[0]) for k in sd if k.startswith(p + ".layers."))
    x = _lin(sd, p + ".proj_in", x)
    for i in range(depth):
        latents = perceiver_attention(sd, f"{p}.layers.{i}.0", x, latents) + latents
        latents = feed_forward(sd, f"{p}.layers.{i}.1", latents) + latents
    return _ln(sd, p + ".norm_out", _lin(sd, p + ".proj_out", latents))


def proj_plus_model(sd, id_embeds, clip_embeds, shortcut=False, scale=1.0, num_tokens=4):
    """ProjPlusModel.forward: id_embeds [B,512], clip_embeds [B,257,1280] -> [B,4,cad]."""
    cad = sd["norm.weight"].shape[0]
    x = _lin(sd, "proj.2", F.gelu(_lin(sd, "proj.0", id_embeds))).reshape(-1, num_tokens, cad)
    x = _ln(sd, "norm", x)
    out = _perceiver_stack(sd, "perceiver_resampler", x, clip_embeds)
    return x + scale * out if shortcut else out


def attention_mlp(sd, x, p="visual_projection"):
    """AttentionMLP.forward: x [n,257,1280] -> [n,1,output_dim] (learned latent token repeated over the batch)."""
    assert p + ".pos_emb.weight" not in sd and not any(k.startswith(p + ".to_latents_from_mean_pooled_seq") for k in sd)
    return _perceiver_stack(sd, p, sd[p + ".latents"].repeat(x.shape[0], 1, 1), x)


def _mlp(sd, p, x, use_residual):
    h = _lin(sd, p + ".fc2", F.gelu(_lin(sd, p + ".fc1", _ln(sd, p + ".layernorm", x))))
    return h + x if use_residual else h


def fuse_module(sd, prompt_embeds, id_embeds, class_tokens_mask, valid_id_mask, p="fuse_module"):
    """prompt_embeds [bs,77,D]; id_embeds [bs,5,1,D]; class_tokens_mask [bs,77] bool; valid_id_mask [bs,5] bool.
    Row r of the valid id embeddings replaces (after fusion) the r-th True position of the flattened token mask."""
    bs, seq, D = prompt_embeds.shape
    ids = id_embeds.to(prompt_embeds.dtype).reshape(-1, id_embeds.shape[-1])[valid_id_mask.flatten()]
    flat = prompt_embeds.reshape(-1, D).clone()
    mask = class_tokens_mask.reshape(-1)
    rows = flat[mask]
    assert rows.shape[0] == ids.shape[0], f"{int(mask.sum())} != {ids.shape[0]}"
    fused = _mlp(sd, p + ".mlp1", torch.cat([rows, ids], dim=-1), False) + rows
    fused = _ln(sd, p + ".layer_norm", _mlp(sd, p + ".mlp2", fused, True))
    flat[mask] = fused.to(flat.dtype)
    return flat.reshape(bs, seq, D)


def facial_encoder(sd, prompt_embeds, multi_image_embeds, class_tokens_mask, valid_id_mask):
    """FacialEncoder.forward: multi_image_embeds [bs,5,257,1280] -> prompt_embeds with the facial rows fused in."""
    bs, n, tl, idim = multi_image_embeds.shape
    id_embeds = attention_mlp(sd, multi_image_embeds.reshape(bs * n, tl, idim)).reshape(bs, n, 1, -1)
    return fuse_module(sd, prompt_embeds, id_embeds, class_tokens_mask, valid_id_mask)


def assemble_prompts(facial_text, uncond_facial_text, text_only, id_tokens, uncond_id_tokens):
    """pipline_StableDiffusion_ConsistentID.py:494-507 -> (null, augmented, text_only) each [1,81,cad], the three prompt
    tensors the denoising loop switches between (oracle/loop_ref.py)."""
    return (torch.cat([uncond_facial_text, uncond_id_tokens], dim=1), torch.cat([facial_text, id_tokens], dim=1),
            torch.cat([text_only, id_tokens], dim=1))
