"""GPU parity of the embedding producers (consistentid_b200/embed.py) against the pinned oracle (oracle/embed_ref.py).

Truth = oracle in fp32 on the CPU; the same oracle run in 16-bit on the GPU gives the error a plain 16-bit evaluation makes
(tests/test_unet_gpu.py criterion: |ours-fp32| <= 3 |eager16-fp32| + 2e-3 max|truth|)."""
import pytest
import torch

from oracle import embed_ref
from tests.test_unet_gpu import _cmp


def _init(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for n, s in shapes.items():
        if len(s) > 1:
            sd[n] = torch.randn(s, generator=g) * (s[-1] ** -0.5)
        elif n.endswith("weight"):
            sd[n] = 1.0 + 0.1 * torch.randn(s, generator=g)
        else:
            sd[n] = 0.1 * torch.randn(s, generator=g)
    return sd


def _to16(sd, dtype):
    return {k: v.cuda().to(dtype) for k, v in sd.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("cad,id_dim,clip_dim,B,n_clip,dtype", [
    (128, 64, 128, 2, 17, torch.float16),        # reduced widths
    (768, 512, 1280, 1, 257, torch.float16),     # SD1.5 widths (pipline_StableDiffusion_ConsistentID.py:89-94)
    (2048, 512, 1280, 2, 257, torch.bfloat16),   # SDXL widths, cond + uncond rows in one call
])
def test_proj_plus_model(cad, id_dim, clip_dim, B, n_clip, dtype):
    from consistentid_b200.embed import ProjPlusModel
    m = ProjPlusModel(cross_attention_dim=cad, id_embeddings_dim=id_dim, clip_embeddings_dim=clip_dim, num_tokens=4, dtype=dtype)
    sd = _init(m.w._shapes, seed=31)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(32)
    idv, clip = torch.randn(B, id_dim, generator=g), torch.randn(B, n_clip, clip_dim, generator=g)
    for shortcut, scale in ((False, 1.0), (True, 0.7)):
        truth = embed_ref.proj_plus_model(sd, idv, clip, shortcut=shortcut, scale=scale)
        eager = embed_ref.proj_plus_model(_to16(sd, dtype), idv.cuda().to(dtype), clip.cuda().to(dtype), shortcut=shortcut, scale=scale)
        out = m(idv.cuda().to(dtype), clip.cuda().to(dtype), shortcut=shortcut, scale=scale)
        torch.cuda.synchronize()
        assert out.shape == (B, 4, cad)
        _cmp(f"ProjPlusModel cad={cad} shortcut={shortcut}", out, truth, eager)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,depth,heads,clip_dim,D,bs,n_clip,dtype", [
    (128, 2, 2, 128, 128, 2, 17, torch.float16),
    (1024, 8, 16, 1280, 768, 1, 257, torch.float16),      # SD1.5 FacialEncoder() defaults (attention.py:73-76)
    (1024, 8, 16, 1280, 2048, 1, 257, torch.bfloat16),    # SDXL: FacialEncoder(..., output_dim=2048, embed_dim=2048)
])
def test_facial_encoder(dim, depth, heads, clip_dim, D, bs, n_clip, dtype):
    from consistentid_b200.embed import FacialEncoder
    m = FacialEncoder(embedding_dim=clip_dim, output_dim=D, embed_dim=D, dtype=dtype, dim=dim, depth=depth, heads=heads)
    sd = _init(m.w._shapes, seed=41)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(42)
    prompt = torch.randn(bs, 77, D, generator=g)
    imgs = torch.randn(bs, 5, n_clip, clip_dim, generator=g)
    for tok_pos, n_valid in (([[5, 9, 20], [1, 76]][:bs], [3, 2][:bs]), ([[]] * bs, [0] * bs)):
        cm, vm = torch.zeros(bs, 77, dtype=torch.bool), torch.zeros(bs, 5, dtype=torch.bool)
        for b in range(bs):
            cm[b, tok_pos[b]] = True
            vm[b, :n_valid[b]] = True
        truth = embed_ref.facial_encoder(sd, prompt, imgs, cm, vm)
        eager = embed_ref.facial_encoder(_to16(sd, dtype), prompt.cuda().to(dtype), imgs.cuda().to(dtype), cm.cuda(), vm.cuda())
        out = m(prompt.cuda().to(dtype), imgs.cuda().to(dtype), cm, vm)
        torch.cuda.synchronize()
        _cmp(f"FacialEncoder dim={dim} D={D} n_valid={n_valid}", out, truth, eager)
        keep = ~cm
        assert torch.equal(out.cpu()[keep], prompt.to(dtype)[keep])                 # rows outside the mask are untouched
    with pytest.raises(AssertionError):                                               # attention.py:44
        cm2 = cm.clone(); cm2[0, 70] = True
        m(prompt.cuda().to(dtype), imgs.cuda().to(dtype), cm2, vm)


@pytest.mark.gpu
def test_producers_feed_the_denoiser_wire_format():
    """[77 fused text rows | 4 id rows] from the producers has the layout B200UNet.set_prompt consumes."""
    from consistentid_b200.embed import FacialEncoder, ProjPlusModel, assemble_prompts
    dtype, cad = torch.float16, 128
    pm = ProjPlusModel(cross_attention_dim=cad, id_embeddings_dim=64, clip_embeddings_dim=128, dtype=dtype)
    pm.load_state_dict(_init(pm.w._shapes, 1))
    fe = FacialEncoder(embedding_dim=128, output_dim=cad, embed_dim=cad, dtype=dtype, dim=128, depth=1, heads=2)
    fe.load_state_dict(_init(fe.w._shapes, 2))
    g = torch.Generator().manual_seed(3)
    r = lambda *s: torch.randn(*s, generator=g).cuda().to(dtype)
    text, neg, text_only = r(1, 77, cad), r(1, 77, cad), r(1, 77, cad)
    cm, vm = torch.zeros(1, 77, dtype=torch.bool), torch.zeros(1, 5, dtype=torch.bool)
    cm[0, [3, 4]] = True; vm[0, :2] = True
    imgs = r(1, 5, 17, 128)
    idt, uidt = pm(r(1, 64), r(1, 17, 128)), pm(torch.zeros(1, 64, device="cuda", dtype=dtype), r(1, 17, 128))
    null, aug, txt = assemble_prompts(fe(text, imgs, cm, vm), fe(neg, torch.zeros_like(imgs), cm, vm), text_only, idt, uidt)
    assert null.shape == aug.shape == txt.shape == (1, 81, cad)
    assert torch.equal(aug[:, 77:], idt) and torch.equal(txt[:, 77:], idt) and torch.equal(null[:, 77:], uidt)
    assert torch.isfinite(aug.float()).all() and torch.isfinite(null.float()).all()
