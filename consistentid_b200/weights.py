"""Weight preparation for the B200 engine: LoRA folding, GEGLU tile interleave, conv weight re-layout,
checkpoint-format handling (adapter_modules positional keys, pipline_StableDiffusion_ConsistentID.py:134-144).
Pure tensor re-arrangement done once at load time (any device)."""
from __future__ import annotations

import torch


def fold_lora(w: torch.Tensor, down: torch.Tensor, up: torch.Tensor, lora_scale: float = 1.0,
              network_alpha=None) -> torch.Tensor:
    """W' = W + lora_scale * (alpha/rank) * up @ down, in fp32, cast back once.
    Exact restatement of ``attn.to_q(x) + lora_scale * to_q_lora(x)`` (attention.py:138, 236) up to 16-bit rounding order."""
    rank = down.shape[0]
    s = lora_scale * ((network_alpha / rank) if network_alpha is not None else 1.0)
    return (w.float() + s * (up.float() @ down.float())).to(w.dtype)


def interleave_geglu(w: torch.Tensor, b: torch.Tensor | None, tile_n: int):
    """Re-order the rows of GEGLU's ``proj`` ([2*inner, K]; first half = value, second = gate) so that each GEMM N-tile of
    width ``tile_n`` holds ``tile_n/2`` value rows followed by the matching ``tile_n/2`` gate rows."""
    inner = w.shape[0] // 2
    half = tile_n // 2
    assert inner % half == 0, (inner, half)
    idx = torch.arange(inner, device=w.device).reshape(-1, half)
    order = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    return w[order].contiguous(), (None if b is None else b[order].contiguous())


def pack_conv3x3(w: torch.Tensor, cin_pad: int | None = None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9 * Cin_pad] with K index = (ky*3 + kx) * Cin_pad + c (zero padded channels)."""
    cout, cin = w.shape[:2]
    cp = cin_pad or cin
    out = torch.zeros((cout, 3, 3, cp), dtype=w.dtype, device=w.device)
    out[..., :cin] = w.permute(0, 2, 3, 1)
    return out.reshape(cout, 9 * cp).contiguous()


def fold_layernorm(w, bias, gamma, beta):
    """LayerNorm(x; gamma, beta) @ w.T + bias  ==  xhat @ (w * gamma).T + (bias + w @ beta)   with xhat = (x - mean) / std.
    Returns (w' in w's 16-bit dtype, bias' in that dtype, colsum fp32 [N]) for the GEMM's folded-LayerNorm epilogue
    (cid_gemm ln_stats / ln_colsum): colsum is taken over the ROUNDED w' so that mean * colsum cancels the mean part of x @ w'.T exactly."""
    w32 = w.float() * gamma.float()[None, :]
    b32 = w.float() @ beta.float()
    if bias is not None:
        b32 = b32 + bias.float()
    wq = w32.to(w.dtype)
    return wq.contiguous(), b32.to(w.dtype).contiguous(), wq.float().sum(dim=1).contiguous()


class TensorIdent:
    """Identity of a tensor's CONTENT, for host-side caches keyed on "the same prompt / weight tensor as last time".

    Keying on ``data_ptr()`` alone is wrong: the reference loop builds its prompt tensor with a fresh ``torch.cat`` every step
    (pipline_StableDiffusion_ConsistentID.py:542-549), the caching allocator recycles freed blocks, and a DIFFERENT prompt can
    land on the SAME address with version 0.  This object therefore holds a strong reference to the tensor that owns the storage
    (so its address cannot be recycled while the cache entry lives) and matches only the same owner object, view geometry and
    version counter.  Inference tensors (no version counter) never match: they are recomputed."""
    __slots__ = ("base", "geom", "version")

    def __init__(self, t):
        self.base = t._base if t._base is not None else t
        self.geom = (t.storage_offset(), tuple(t.shape), tuple(t.stride()), t.dtype)
        try:
            self.version = t._version
        except RuntimeError:
            self.version = None

    def matches(self, t) -> bool:
        if t is None or self.version is None:
            return False
        try:
            v = t._version
        except RuntimeError:
            return False
        base = t._base if t._base is not None else t
        return base is self.base and v == self.version and (t.storage_offset(), tuple(t.shape), tuple(t.stride()), t.dtype) == self.geom


def same_tensors(idents, tensors) -> bool:
    """idents: tuple of TensorIdent | None recorded earlier; tensors: the tensors (or None) of this call."""
    if idents is None or len(idents) != len(tensors):
        return False
    for i, t in zip(idents, tensors):
        if (i is None) != (t is None):
            return False
        if i is not None and not i.matches(t):
            return False
    return True
