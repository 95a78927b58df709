"""Generate golden vectors for the two attention processors FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference): imports the reference's ``attention.py`` verbatim through the
2-symbol ``oracle/diffusers_shim`` (diffusers is not installable here) and records, for seeded inputs, the outputs of
``Consistent_AttProcessor.__call__`` (attention.py:110-174) and ``Consistent_IPAttProcessor.__call__`` (attention.py:207-294)
in fp32 on CPU.  The fixtures travel to the GPU box; /root/reference does not.

    python tests/golden/make_golden.py
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
from oracle.unet_ref import Attention  # noqa: E402  (diffusers Attention stand-in: fields used at attention.py:120-172)


def make_case(seed, C, heads, cad, N, B, rank, scale):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    attn1, attn2 = Attention(C, None, heads, C // heads), Attention(C, cad, heads, C // heads)
    p1 = ref_attention.Consistent_AttProcessor(hidden_size=C, cross_attention_dim=None, rank=rank)
    p2 = ref_attention.Consistent_IPAttProcessor(hidden_size=C, cross_attention_dim=cad, rank=rank, scale=scale, num_tokens=4)
    for m in (attn1, attn2, p1, p2):
        for p in m.parameters():
            p.data = rnd(*p.shape, std=(p.shape[-1] ** -0.5 if p.ndim > 1 else 0.1))
    x = rnd(B, N, C)
    ehs = rnd(B, 81, cad)
    with torch.no_grad():
        y1 = p1(attn1, x)
        y2 = p2(attn2, x, encoder_hidden_states=ehs)
    return dict(meta=dict(seed=seed, C=C, heads=heads, cad=cad, N=N, B=B, rank=rank, scale=scale),
                attn1=attn1.state_dict(), attn2=attn2.state_dict(), proc1=p1.state_dict(), proc2=p2.state_dict(),
                x=x, ehs=ehs, y_self=y1, y_cross=y2)


if __name__ == "__main__":
    cases = [make_case(11, 64, 2, 64, 64, 2, 8, 1.0),      # d=32
             make_case(12, 128, 2, 128, 128, 1, 16, 0.6),  # d=64
             make_case(13, 320, 8, 64, 64, 1, 8, 1.0)]     # d=40 (SD1.5 level-0 head dim)
    out = os.path.join(HERE, "processors_golden.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
