"""GPU parity of the CLIP ViT image encoder (consistentid_b200/clip.py) against oracle/clip_ref.py (pinned on the installed
transformers implementation, tests/test_clip_cpu.py): reduced widths, and the real ViT-H/14 geometry the reference loads
(1280 wide, 16 heads of 80, 32 layers, MLP 5120, 224 px -> 257 tokens) on a batch of 12 crops (face + 5 facial regions + their zero images,
pipline_StableDiffusion_ConsistentID.py:182-183, 202-203)."""
import pytest
import torch

from oracle import clip_ref
from tests.test_unet_gpu import _cmp


def _weights(C, heads, layers, inter, image, patch, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    n_tok = (image // patch) ** 2 + 1
    sd = {"vision_model.embeddings.patch_embedding.weight": r(C, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
          "vision_model.embeddings.class_embedding": r(C, std=0.5), "vision_model.embeddings.position_embedding.weight": r(n_tok, C, std=0.1),
          "vision_model.pre_layrnorm.weight": 1 + r(C, std=0.1), "vision_model.pre_layrnorm.bias": r(C, std=0.1)}
    for i in range(layers):
        b = f"vision_model.encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            sd[b + ln + ".weight"], sd[b + ln + ".bias"] = 1 + r(C, std=0.1), r(C, std=0.1)
        for q in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[b + f"self_attn.{q}.weight"], sd[b + f"self_attn.{q}.bias"] = r(C, C, std=C ** -0.5), r(C, std=0.1)
        sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"] = r(inter, C, std=C ** -0.5), r(inter, std=0.1)
        sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"] = r(C, inter, std=inter ** -0.5), r(C, std=0.1)
    return sd


@pytest.mark.gpu
@pytest.mark.parametrize("C,heads,layers,inter,image,batch,dtype", [
    (128, 2, 3, 256, 42, 2, torch.float16), (192, 2, 4, 640, 56, 2, torch.bfloat16),
    (1280, 16, 32, 5120, 224, 12, torch.float16)])          # laion ViT-H/14, the encoder the reference loads
def test_clip_vision_encoder_parity(C, heads, layers, inter, image, batch, dtype):
    from consistentid_b200.clip import B200CLIPVisionEncoder
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = _weights(C, heads, layers, inter, image, 14)
    x = torch.randn(batch, 3, image, image, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        sd32 = {k: v.cuda() for k, v in sd.items()}                         # fp32 truth evaluated on the GPU (TF32 off)
        truth = clip_ref.penultimate_hidden_state(sd32, x.cuda(), heads).cpu()
        del sd32
        sd16 = {k: v.cuda().to(dtype) for k, v in sd.items()}
        eager = clip_ref.penultimate_hidden_state(sd16, x.cuda().to(dtype), heads)
        del sd16
    enc = B200CLIPVisionEncoder(sd, num_attention_heads=heads, dtype=dtype)
    out = enc(x.cuda().to(dtype))
    torch.cuda.synchronize()
    assert out.shape == truth.shape
    _cmp(f"clip vision C={C} layers={layers} batch={batch} {dtype}", out, truth, eager)
