"""Generate golden vectors for the embedding producers FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference): imports the reference's ``functions.py`` and ``attention.py`` verbatim
(the latter through the 2-symbol ``oracle/diffusers_shim``) and records, for seeded weights/inputs at reduced widths, the
fp32 CPU outputs of ``ProjPlusModel.forward`` (functions.py:520-528), ``AttentionMLP.forward`` (functions.py:571-592),
``FuseModule.forward`` (attention.py:25-48) and ``FacialEncoder.forward`` (attention.py:78-88).  The reference classes fix the
perceiver width of FacialEncoder at 1024 x 8 layers (~100 M parameters); to keep the fixture small its ``visual_projection`` /
``fuse_module`` attributes are replaced by narrower instances OF THE SAME REFERENCE CLASSES - the forward code is untouched.

    python tests/golden/make_embed_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
sys.path.insert(0, "/root/reference")

import attention as ref_attention  # noqa: E402  (the reference, verbatim)
import functions as ref_functions  # noqa: E402  (the reference, verbatim)


def _reinit(module, g):
    for n, p in module.named_parameters():
        if p.ndim > 1:
            p.data = torch.randn(p.shape, generator=g) * (p.shape[-1] ** -0.5)
        elif n.endswith("weight"):                   # every 1-D weight in these modules is a LayerNorm gain
            p.data = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        else:
            p.data = 0.1 * torch.randn(p.shape, generator=g)


def proj_case(seed, cad, id_dim, clip_dim, B, n_clip):
    g = torch.Generator().manual_seed(seed)
    m = ref_functions.ProjPlusModel(cross_attention_dim=cad, id_embeddings_dim=id_dim, clip_embeddings_dim=clip_dim, num_tokens=4).eval()
    _reinit(m, g)
    idv, clip = torch.randn(B, id_dim, generator=g), torch.randn(B, n_clip, clip_dim, generator=g)
    with torch.no_grad():
        y, y_sc = m(idv, clip), m(idv, clip, shortcut=True, scale=0.7)
    return dict(kind="proj_plus", meta=dict(seed=seed, cad=cad, id_dim=id_dim, clip_dim=clip_dim), sd=m.state_dict(), id_embeds=idv, clip_embeds=clip,
                y=y, y_shortcut=y_sc, shortcut_scale=0.7)


def facial_case(seed, dim, depth, clip_dim, D, bs, n_clip, masks):
    g = torch.Generator().manual_seed(seed)
    enc = ref_attention.FacialEncoder(embedding_dim=clip_dim, output_dim=D, embed_dim=D)
    enc.visual_projection = ref_functions.AttentionMLP(dim=dim, depth=depth, dim_head=64, heads=dim // 64, embedding_dim=clip_dim, output_dim=D)
    enc.fuse_module = ref_attention.FuseModule(D)
    enc.eval()
    _reinit(enc, g)
    prompt = torch.randn(bs, 77, D, generator=g)
    imgs = torch.randn(bs, 5, n_clip, clip_dim, generator=g)
    out = dict(kind="facial_encoder", meta=dict(seed=seed, dim=dim, depth=depth, clip_dim=clip_dim, D=D), sd=enc.state_dict(), prompt_embeds=prompt,
               multi_image_embeds=imgs, cases=[])
    with torch.no_grad():
        out["visual_projection"] = enc.visual_projection(imgs.reshape(bs * 5, n_clip, clip_dim))
        for tok_pos, n_valid in masks:
            cm = torch.zeros(bs, 77, dtype=torch.bool)
            vm = torch.zeros(bs, 5, dtype=torch.bool)
            for b in range(bs):
                cm[b, tok_pos[b]] = True
                vm[b, :n_valid[b]] = True
            y = enc(prompt.clone(), imgs, cm, vm)
            out["cases"].append(dict(class_tokens_mask=cm, valid_id_mask=vm, y=y))
    return out


if __name__ == "__main__":
    cases = [proj_case(21, 128, 32, 96, 2, 17),
             proj_case(22, 64, 64, 64, 1, 9),
             facial_case(23, 128, 2, 96, 128, 1, 17, [([[5, 9, 20]], [3]), ([[]], [0]), ([[1, 2, 3, 40, 76]], [5])]),
             facial_case(24, 64, 1, 64, 64, 2, 9, [([[4, 8], [10]], [2, 1])])]
    out = os.path.join(HERE, "embed_golden.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
