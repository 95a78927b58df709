"""CPU oracle: the CLIP ViT image encoder in front of the embedding producers,
``self.image_encoder(clip_image, output_hidden_states=True).hidden_states[-2]``
(pipline_StableDiffusion_ConsistentID.py:182-183, 202-203; model ``laion/CLIP-ViT-H-14-laion2B-s32B-b79K``: hidden 1280, 32 layers,
16 heads, MLP 5120, patch 14, image 224 -> 257 tokens, GELU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The algorithm lives in the third-party ``transformers`` (reference pin
``transformers==4.36``-era ``CLIPVisionModelWithProjection``); an implementation of it IS importable in this image, so this restatement is
PINNED: tests/test_clip_cpu.py checks it against ``transformers.CLIPVisionModelWithProjection`` on random weights (all hidden states), and the
parameter names are that model's ``state_dict`` keys.

hidden_states[0] = pre_layrnorm(embeddings); hidden_states[i] = output of encoder layer i.  ``hidden_states[-2]`` therefore SKIPS the last
encoder layer and the post_layernorm.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def embeddings(sd, pixel_values, p="vision_model.embeddings"):
    w = sd[p + ".patch_embedding.weight"]                                  # [C, 3, P, P], stride P, no bias
    x = F.conv2d(pixel_values.to(w.dtype), w, stride=w.shape[-1]).flatten(2).transpose(1, 2)       # [B, n_patches, C]
    cls = sd[p + ".class_embedding"].expand(x.shape[0], 1, -1)
    return torch.cat([cls, x], dim=1) + sd[p + ".position_embedding.weight"][None]


def encoder_layer(sd, p, x, heads, act="gelu"):
    b, n, c = x.shape
    d = c // heads
    h = _ln(sd, p + ".layer_norm1", x)
    split = lambda t: t.reshape(b, n, heads, d).transpose(1, 2)
    q, k, v = split(_lin(sd, p + ".self_attn.q_proj", h)), split(_lin(sd, p + ".self_attn.k_proj", h)), split(_lin(sd, p + ".self_attn.v_proj", h))
    w = torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2), dim=-1)
    x = x + _lin(sd, p + ".self_attn.out_proj", (w @ v).transpose(1, 2).reshape(b, n, c))
    h = _lin(sd, p + ".mlp.fc1", _ln(sd, p + ".layer_norm2", x))
    h = F.gelu(h) if act == "gelu" else h * torch.sigmoid(1.702 * h)      # "quick_gelu" of the OpenAI checkpoints
    return x + _lin(sd, p + ".mlp.fc2", h)


def hidden_states(sd, pixel_values, heads, act="gelu"):
    """All hidden states, as ``output_hidden_states=True`` returns them (len = layers + 1)."""
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("vision_model.encoder.layers."))
    hs = [_ln(sd, "vision_model.pre_layrnorm", embeddings(sd, pixel_values))]
    for i in range(n_layers):
        hs.append(encoder_layer(sd, f"vision_model.encoder.layers.{i}", hs[-1], heads, act))
    return hs


def penultimate_hidden_state(sd, pixel_values, heads, act="gelu"):
    """What the reference feeds to ProjPlusModel / FacialEncoder: ``hidden_states[-2]`` [B, 257, 1280] (the last layer is never needed)."""
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("vision_model.encoder.layers."))
    x = _ln(sd, "vision_model.pre_layrnorm", embeddings(sd, pixel_values))
    for i in range(n_layers - 1):
        x = encoder_layer(sd, f"vision_model.encoder.layers.{i}", x, heads, act)
    return x
