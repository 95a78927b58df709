"""oracle/clip_ref.py (CLIP ViT image encoder) pinned against the transformers implementation installed in this image."""
import pytest
import torch

from oracle import clip_ref

transformers = pytest.importorskip("transformers")


@pytest.mark.parametrize("act,heads,layers", [("gelu", 2, 3), ("quick_gelu", 4, 2)])
def test_clip_vision_oracle_matches_transformers(act, heads, layers):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=160, num_hidden_layers=layers, num_attention_heads=heads, image_size=42, patch_size=14,
                           projection_dim=32, hidden_act=act)
    m = CLIPVisionModelWithProjection(cfg).eval()
    for p in m.parameters():                                   # default init leaves biases at zero: make every term count
        if p.ndim == 1:
            p.data += 0.1 * torch.randn_like(p)
    x = torch.randn(2, 3, 42, 42)
    with torch.no_grad():
        want = m(x, output_hidden_states=True).hidden_states
    sd = m.state_dict()
    got = clip_ref.hidden_states(sd, x, heads, act)
    assert len(got) == len(want) == layers + 1 and got[0].shape == (2, 10, 64)          # 3x3 patches + class token
    for g, w in zip(got, want):
        assert torch.allclose(g, w, atol=2e-5, rtol=2e-5), (g - w).abs().max()
    assert torch.allclose(clip_ref.penultimate_hidden_state(sd, x, heads, act), want[-2], atol=2e-5, rtol=2e-5)


def test_vit_h_14_geometry():
    """laion/CLIP-ViT-H-14: 224 / 14 = 16 -> 256 patches + class token = the 257 rows ProjPlusModel / FacialEncoder consume (functions.py:571, attention.py:80)."""
    sd = {"vision_model.embeddings.patch_embedding.weight": torch.zeros(8, 3, 14, 14), "vision_model.embeddings.class_embedding": torch.zeros(8),
          "vision_model.embeddings.position_embedding.weight": torch.zeros(257, 8)}
    assert clip_ref.embeddings(sd, torch.zeros(1, 3, 224, 224)).shape == (1, 257, 8)
