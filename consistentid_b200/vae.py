"""VAE decode on the GPU: the step right after the denoising loop,
``image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]``
(pipline_StableDiffusion_ConsistentID.py:586; SDXL pipline_StableDiffusionXL_ConsistentID.py:669-684).

``B200VAEDecoder`` exposes what the pipelines touch on ``self.vae`` in that direction - ``.config.scaling_factor``, ``.dtype``, ``.device``,
``.decode(z, return_dict=False)[0]`` - and takes the ``decoder.*`` / ``post_quant_conv.*`` entries of a diffusers ``AutoencoderKL``
state_dict (names as in diffusers 0.23; the pre-0.20 attention names query/key/value/proj_attn are accepted too).

It reuses the hot path's kernels on NHWC rows: implicit-GEMM ``cid_conv3x3`` (conv_in with the 4 latent channels zero-padded to one
128-byte row, 3-channel conv_out on the 16-wide tile), ``cid_gn_stats``/``cid_gn_apply`` (eps 1e-6, SiLU fused), ``cid_upsample2x``,
``cid_gemm`` for the 1x1 ``post_quant_conv`` / shortcuts / attention projections.  The mid-block attention is single-head with d = 512 -
beyond the flash kernel's 160 - so it runs as ``Q.K^T`` (scale fused in fp32 before the 16-bit store, which is also what makes the
reference's fp32 upcast of the SDXL VAE unnecessary here) -> ``cid_softmax_rows`` -> ``P.V`` per image.  16-bit CUDA tensors only.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .lib import EPI_QKV
from .weights import pack_conv3x3

CIN_PAD = 64
_OLD_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


class B200VAEDecoder:
    def __init__(self, state_dict, scaling_factor=0.18215, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, latent_channels=4, out_channels=3, dtype=torch.float16, device="cuda"):
        self.dtype, self.device = dtype, torch.device(device)
        # AutoencoderKL config fields the pipelines read.  force_upcast is False: the SDXL pipeline's
        # ``vae.dtype == float16 and vae.config.force_upcast`` branch (pipline_StableDiffusionXL_ConsistentID.py:669-675) must not try to
        # ``.to(float32)`` this engine.  The fp32 score scaling here only removes the ATTENTION overflow; the real SDXL VAE weights also
        # overflow fp16 in the resnet / upsample activations (that is why diffusers sets force_upcast), so decode SDXL latents with
        # dtype=torch.bfloat16 (same exponent range as fp32) - parity has only been shown on synthetic weights.
        self.config = SimpleNamespace(scaling_factor=scaling_factor, block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, latent_channels=latent_channels, out_channels=out_channels, in_channels=out_channels,
                                      force_upcast=False, act_fn="silu", sample_size=512)
        self.groups = norm_num_groups
        ops.ensure_workspace(self.device)
        sd = {}
        for k, v in state_dict.items():
            for old, new in _OLD_ATTN.items():
                k = k.replace(f".attentions.0.{old}.", f".attentions.0.{new}.")
            sd[k] = v
        W = lambda n: sd[n].detach().to(device=self.device, dtype=dtype).contiguous()
        boc = tuple(block_out_channels)
        top = boc[-1]
        self.p = p = {}

        def need(name, shape):
            if name not in sd or tuple(sd[name].shape) != tuple(shape):
                raise KeyError(f"vae state_dict: missing/mis-shaped '{name}' (want {tuple(shape)})")

        def conv(name, cout, cin, cin_pad=None):
            need(name + ".weight", (cout, cin, 3, 3)); need(name + ".bias", (cout,))
            p[name + ".w"], p[name + ".b"] = pack_conv3x3(W(name + ".weight"), cin_pad), W(name + ".bias")

        def norm(name, c):
            need(name + ".weight", (c,)); need(name + ".bias", (c,))
            p[name + ".g"], p[name + ".b"] = W(name + ".weight"), W(name + ".bias")

        def resnet(name, cin, cout):
            norm(name + ".norm1", cin); conv(name + ".conv1", cout, cin); norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout)
            if cin != cout:
                need(name + ".conv_shortcut.weight", (cout, cin, 1, 1))
                p[name + ".sc.w"], p[name + ".sc.b"] = W(name + ".conv_shortcut.weight").reshape(cout, cin).contiguous(), W(name + ".conv_shortcut.bias")

        need("post_quant_conv.weight", (latent_channels, latent_channels, 1, 1))
        wpq = torch.zeros((latent_channels, CIN_PAD), dtype=dtype, device=self.device)
        wpq[:, :latent_channels] = W("post_quant_conv.weight").reshape(latent_channels, latent_channels)
        p["pq.w"], p["pq.b"] = wpq, W("post_quant_conv.bias")
        conv("decoder.conv_in", top, latent_channels, CIN_PAD)
        resnet("decoder.mid_block.resnets.0", top, top); resnet("decoder.mid_block.resnets.1", top, top)
        a = "decoder.mid_block.attentions.0"
        norm(a + ".group_norm", top)
        for q in ("to_q", "to_k", "to_v", "to_out.0"):
            need(f"{a}.{q}.weight", (top, top)) if sd[f"{a}.{q}.weight"].ndim == 2 else need(f"{a}.{q}.weight", (top, top, 1, 1))
        lin = lambda n: W(n).reshape(top, top).contiguous()
        p[a + ".qkv.w"] = torch.cat([lin(f"{a}.to_q.weight"), lin(f"{a}.to_k.weight"), lin(f"{a}.to_v.weight")], 0).contiguous()
        p[a + ".qkv.b"] = torch.cat([W(f"{a}.to_q.bias"), W(f"{a}.to_k.bias"), W(f"{a}.to_v.bias")], 0).contiguous()
        p[a + ".o.w"], p[a + ".o.b"] = lin(f"{a}.to_out.0.weight"), W(f"{a}.to_out.0.bias")
        self.blocks = []
        prev = top
        rev = list(reversed(boc))
        for i, ch in enumerate(rev):
            names = []
            for j in range(layers_per_block + 1):
                n = f"decoder.up_blocks.{i}.resnets.{j}"
                resnet(n, prev if j == 0 else ch, ch)
                names.append((n, prev if j == 0 else ch, ch))
            up = None
            if i != len(rev) - 1:
                up = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                conv(up, ch, ch)
            self.blocks.append((names, up, ch))
            prev = ch
        norm("decoder.conv_norm_out", boc[0])
        conv("decoder.conv_out", out_channels, boc[0])
        self._inv_scale = torch.tensor([1.0 / scaling_factor], dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ building blocks on NHWC rows
    def _new(self, *shape):
        return torch.empty(shape, dtype=self.dtype, device=self.device)

    def _gn(self, x, C, NB, HW, name, silu):
        sums = torch.empty((NB, self.groups, 2), dtype=torch.float32, device=self.device)
        ops.gn_stats(x, C, None, 0, NB, HW, self.groups, sums)
        return ops.gn_apply(x, C, None, 0, NB, HW, self.groups, sums, self.p[name + ".g"], self.p[name + ".b"], 1e-6, silu, self._new(NB * HW, C))

    def _resnet(self, x, name, cin, cout, NB, h, w):
        p, HW = self.p, h * w
        a1 = self._gn(x, cin, NB, HW, name + ".norm1", True)
        h1 = ops.conv3x3(a1, p[name + ".conv1.w"], self._new(NB * HW, cout), NB, h, w, cin, cout, bias=p[name + ".conv1.b"])
        a2 = self._gn(h1, cout, NB, HW, name + ".norm2", True)
        sc = x if cin == cout else ops.gemm(x, p[name + ".sc.w"], self._new(NB * HW, cout), bias=p[name + ".sc.b"])
        return ops.conv3x3(a2, p[name + ".conv2.w"], self._new(NB * HW, cout), NB, h, w, cout, cout, bias=p[name + ".conv2.b"], residual=sc)

    def _attention(self, x, C, NB, N):
        p, a = self.p, "decoder.mid_block.attentions.0"
        if N % 64:
            raise ValueError(f"VAE mid-block attention needs H*W % 64 == 0 (got {N})")
        t = self._gn(x, C, NB, N, a + ".group_norm", False)
        qk, vt = self._new(NB * N, 2 * C), self._new(NB, C, N)
        ops.gemm(t, p[a + ".qkv.w"], qk, bias=p[a + ".qkv.b"], epi=EPI_QKV, vt=vt, n_split=2 * C, heads=1, hdim=C, ntok=N)
        k_all = qk[:, C:].contiguous()                       # the B operand of Q.K^T must be a dense [N, C] matrix
        o, s = self._new(NB * N, C), self._new(N, N)
        for b in range(NB):
            ops.gemm(qk[b * N:(b + 1) * N, :C], k_all[b * N:(b + 1) * N], s, out_scale=C ** -0.5)
            ops.softmax_rows(s, N, N)
            ops.gemm(s, vt[b], o[b * N:(b + 1) * N])
        return ops.gemm(o, p[a + ".o.w"], self._new(NB * N, C), bias=p[a + ".o.b"], residual=x)

    # ------------------------------------------------------------------ diffusers-facing surface
    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None, _scale_dev=None):
        if not (z.is_cuda and z.dtype == self.dtype and z.ndim == 4 and z.shape[1] == self.config.latent_channels):
            raise TypeError(f"B200VAEDecoder.decode: expected a CUDA {self.dtype} tensor [B, {self.config.latent_channels}, h, w] (no CPU/fp32 path)")
        p = self.p
        NB, L, h, w = z.shape
        M = NB * h * w
        zin = ops.nchw_to_nhwc_pad(z.contiguous(), self._new(M, CIN_PAD), NB, L, h * w, CIN_PAD, scale_dev=_scale_dev)
        pq = torch.zeros((M, CIN_PAD), dtype=self.dtype, device=self.device)
        ops.gemm(zin, p["pq.w"], pq[:, :L], bias=p["pq.b"])                                       # post_quant_conv (1x1)
        top = self.config.block_out_channels[-1]
        x = ops.conv3x3(pq, p["decoder.conv_in.w"], self._new(M, top), NB, h, w, CIN_PAD, top, bias=p["decoder.conv_in.b"])
        x = self._resnet(x, "decoder.mid_block.resnets.0", top, top, NB, h, w)
        x = self._attention(x, top, NB, h * w)
        x = self._resnet(x, "decoder.mid_block.resnets.1", top, top, NB, h, w)
        for names, up, ch in self.blocks:
            for n, cin, cout in names:
                x = self._resnet(x, n, cin, cout, NB, h, w)
            if up is not None:
                big = ops.upsample2x(x, self._new(NB * 4 * h * w, ch), NB, h, w, ch)
                h, w = 2 * h, 2 * w
                x = ops.conv3x3(big, p[up + ".w"], self._new(NB * h * w, ch), NB, h, w, ch, ch, bias=p[up + ".b"])
        c0, oc = self.config.block_out_channels[0], self.config.out_channels
        act = self._gn(x, c0, NB, h * w, "decoder.conv_norm_out", True)
        rows = ops.conv3x3(act, p["decoder.conv_out.w"], self._new(NB * h * w, 4), NB, h, w, c0, oc, bias=p["decoder.conv_out.b"])
        img = ops.rows_to_nchw(rows, 4, torch.empty((NB, oc, h, w), dtype=self.dtype, device=self.device), NB, oc, h * w)
        return SimpleNamespace(sample=img) if return_dict else (img,)

    def decode_latents(self, latents):
        """``decode(latents / scaling_factor)`` with the division folded into the NCHW -> NHWC conversion kernel."""
        return self.decode(latents, return_dict=False, _scale_dev=self._inv_scale)[0]
