"""Torch (CPU, fp32-math) stand-ins for the libcidb200 entry points, with the same argument conventions as ``consistentid_b200.ops``.

TEST INFRASTRUCTURE: lets the ``-m "not gpu"`` suite drive the HOST logic of the engines (weight packing and LoRA folding, buffer reuse, the
launch sequence of the UNet / ControlNet / denoising loop / VAE / embedding producers, scheduler coefficient tables) against the oracle
without a GPU.  Each function restates what the kernel of the same name computes (see include/cidb200.h); the kernels themselves are
checked against PyTorch on the GPU by tests/kernel_checks.py.  ``install(monkeypatch)`` swaps them into ``consistentid_b200.ops`` and makes
tensors report ``is_cuda`` so the engines' device checks pass.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from consistentid_b200 import lib
from consistentid_b200.lib import EPI_GEGLU, EPI_QKV


def _f(t):
    return None if t is None else t.float()


def _add_chan_stats(chan_stats, y, rows_per_sample):
    if chan_stats is not None:
        M, N = y.shape
        yy = y.float().view(M // rows_per_sample, rows_per_sample, N)
        chan_stats.view(-1, N, 2).add_(torch.stack([yy.sum(1), (yy * yy).sum(1)], dim=-1))


def gemm(a, w, out, bias=None, residual=None, rowbias=None, rows_per_group=1, a2=None, epi=0, vt=None, n_split=0, heads=0, hdim=0, ntok=0,
         out_scale=1.0, chan_stats=None, stats_rows=0, row_stats=None, ln=None):
    x = a.float() if a2 is None else torch.cat([a.float(), a2.float()], dim=1)
    y = x @ w.float().T
    M, N = y.shape
    if ln is not None:            # folded LayerNorm: rstd * (acc - mean * colsum), then the bias (which carries W . beta)
        st, cs, eps = ln
        K = x.shape[1]
        mean = st[:, 0] / K
        rstd = torch.rsqrt((st[:, 1] / K - mean * mean).clamp_min(0) + eps)
        y = rstd[:, None] * (y - mean[:, None] * cs.float()[None, :])
    if bias is not None:
        y = y + bias.float()
    if rowbias is not None:
        y = y + rowbias.float().repeat_interleave(rows_per_group, dim=0)[:M]
    if epi == EPI_GEGLU:
        tile = lib.gemm_tile_n(N, EPI_GEGLU)
        y = y.view(M, N // tile, tile)
        out.copy_((y[..., :tile // 2] * F.gelu(y[..., tile // 2:])).reshape(M, N // 2))
        return out
    if epi == EPI_QKV:
        out.copy_(y[:, :n_split])
        B = M // ntok
        vt.copy_(y[:, n_split:].reshape(B, ntok, heads, hdim).permute(0, 2, 3, 1).reshape(B * heads, hdim, ntok))
        return out
    if residual is not None:
        y = y + residual.float()
    y = y * out_scale
    if epi == 3:          # EPI_GELU
        y = F.gelu(y)
    out.copy_(y)
    _add_chan_stats(chan_stats, y, stats_rows)
    if row_stats is not None:
        row_stats.add_(torch.stack([y.sum(1), (y * y).sum(1)], dim=1))
    return out


def conv3x3(x, w, out, NB, H, W, Cin, Cout, bias=None, residual=None, rowbias=None, stride2=False, out_scale=1.0, chan_stats=None):
    wt = w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)                         # [Cout, 9*Cin] is (ky, kx, c) order
    if stride2:                                                                      # x: phase-split copy [NB, 4, H, W, Cin] of [NB, 2H, 2W, Cin]
        ps = x.float().view(NB, 2, 2, H, W, Cin)
        full = torch.zeros(NB, 2 * H, 2 * W, Cin)
        for py in range(2):
            for px in range(2):
                full[:, py::2, px::2] = ps[:, py, px]
        y = F.conv2d(full.permute(0, 3, 1, 2), wt, stride=2, padding=1)
    else:
        y = F.conv2d(x.float().view(NB, H, W, Cin).permute(0, 3, 1, 2), wt, padding=1)
    y = y.permute(0, 2, 3, 1).reshape(NB * H * W, Cout)
    if bias is not None:
        y = y + bias.float()
    if rowbias is not None:
        y = y + rowbias.float().repeat_interleave(H * W, dim=0)
    if residual is not None:
        y = y + residual.float()[:, :Cout]
    out[:, :Cout] = y * out_scale
    _add_chan_stats(chan_stats, y * out_scale, H * W)
    return out


def _heads(t, B, N, H, d):
    return t.float().reshape(B, N, -1)[..., :H * d].reshape(B, N, H, d).transpose(1, 2)


def attn_self(q, k, vt, out, B, H, N, d, n_valid=None):
    nv = N if n_valid is None else n_valid
    v = vt.float().view(B, H, d, N).transpose(2, 3)
    o = F.scaled_dot_product_attention(_heads(q, B, N, H, d), _heads(k, B, N, H, d)[:, :, :nv], v[:, :, :nv])
    out.copy_(o.transpose(1, 2).reshape(B * N, H * d))
    return out


def pack_cross_kv(k_text, v_text, k_ip, v_ip, k_cat, vt_cat, B, C, heads, n_text, n_ip):
    d = C // heads
    k_cat.zero_(); vt_cat.zero_()
    kc, vc = k_cat.view(B, 96, C), torch.zeros(B, 96, C, dtype=vt_cat.dtype)
    kc[:, :n_text] = k_text.view(B, n_text, C); vc[:, :n_text] = v_text.view(B, n_text, C)
    if n_ip:
        kc[:, 80:80 + n_ip] = k_ip.view(B, n_ip, C); vc[:, 80:80 + n_ip] = v_ip.view(B, n_ip, C)
    vt_cat.copy_(vc.view(B, 96, heads, d).permute(0, 2, 3, 1).reshape(B * heads, d, 96))


def attn_cross(q, k_cat, vt_cat, out, B, H, N, d, n_text, n_ip, ip_scale):
    qh = _heads(q, B, N, H, d)
    k = k_cat.float().view(B, 96, H, d).transpose(1, 2)
    v = vt_cat.float().view(B, H, d, 96).transpose(2, 3)
    rnd = lambda t: t.to(out.dtype).float()                                          # each branch is rounded to the storage type first
    o = rnd(F.scaled_dot_product_attention(qh, k[:, :, :n_text], v[:, :, :n_text]))
    if n_ip:
        o = o + ip_scale * rnd(F.scaled_dot_product_attention(qh, k[:, :, 80:80 + n_ip], v[:, :, 80:80 + n_ip]))
    out.copy_(o.transpose(1, 2).reshape(B * N, H * d))
    return out


def gn_stats(x1, C1, x2, C2, NB, HW, groups, sums, zero_sums=True):
    x = x1.float().reshape(NB, HW, C1)
    if C2:
        x = torch.cat([x, x2.float().reshape(NB, HW, C2)], dim=-1)
    g = x.reshape(NB, HW, groups, (C1 + C2) // groups)
    if zero_sums:
        sums.zero_()
    sums[..., 0] += g.sum(dim=(1, 3))
    sums[..., 1] += (g * g).sum(dim=(1, 3))


def gn_apply(x1, C1, x2, C2, NB, HW, groups, sums, gamma, beta, eps, silu, out, zero_next=None):
    C = C1 + C2
    x = x1.float().reshape(NB, HW, C1)
    if C2:
        x = torch.cat([x, x2.float().reshape(NB, HW, C2)], dim=-1)
    n = HW * (C // groups)
    mean = sums[..., 0] / n
    rstd = torch.rsqrt((sums[..., 1] / n - mean * mean).clamp_min(0) + eps)
    y = (x.reshape(NB, HW, groups, C // groups) - mean[:, None, :, None]) * rstd[:, None, :, None]
    y = y.reshape(NB, HW, C) * gamma.float() + beta.float()
    out.copy_((F.silu(y) if silu else y).reshape(NB * HW, C))
    if zero_next is not None:
        zero_next.zero_()
    return out


def gn_apply_ch(x1, C1, sums1, x2, C2, sums2, NB, HW, groups, gamma, beta, eps, silu, out):
    ch = sums1.view(NB, C1, 2)
    if C2:
        ch = torch.cat([ch, sums2.view(NB, C2, 2)], dim=1)
    sums = ch.view(NB, groups, (C1 + C2) // groups, 2).sum(2)
    return gn_apply(x1, C1, x2, C2, NB, HW, groups, sums, gamma, beta, eps, silu, out)


def gn_small_ok(C1, C2, HW, groups):
    C = C1 + C2
    cpg = C // groups
    return C % groups == 0 and cpg % 8 == 0 and C1 % cpg == 0 and HW * cpg <= 32768


def gn_small(x1, C1, x2, C2, NB, HW, groups, gamma, beta, eps, silu, out):
    assert gn_small_ok(C1, C2, HW, groups)
    x = x1.float().reshape(NB, HW, C1)
    if C2:
        x = torch.cat([x, x2.float().reshape(NB, HW, C2)], -1)
    yv = F.group_norm(x.permute(0, 2, 1), groups, gamma.float(), beta.float(), eps)
    if silu:
        yv = F.silu(yv)
    out.copy_(yv.permute(0, 2, 1).reshape(out.shape))
    return out


def layernorm(x, gamma, beta, out, rows, C, eps=1e-5):
    out.copy_(F.layer_norm(x.float().reshape(rows, C), (C,), gamma.float(), beta.float(), eps).reshape(out.shape))
    return out


def layernorm_rows(x, gamma, beta, out, rows, C, eps=1e-5, rows_per_group=None, x_group_rows=None, x_row0=0, y_group_rows=None, y_row0=0):
    rpg = rows if rows_per_group is None else rows_per_group
    xg, yg = (rpg if x_group_rows is None else x_group_rows), (rpg if y_group_rows is None else y_group_rows)
    for r in range(rows):
        g, i = divmod(r, rpg)
        out[g * yg + y_row0 + i, :C] = F.layer_norm(x[g * xg + x_row0 + i, :C].float(), (C,), gamma.float(), beta.float(), eps)
    return out


def upsample2x(x, out, NB, H, W, C):
    out.copy_(x.view(NB, H, W, C).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(out.shape))
    return out


def phase_split(x, out, NB, H, W, C):
    xs = x.view(NB, H, W, C)
    out.copy_(torch.stack([xs[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1).reshape(out.shape))
    return out


def nchw_to_nhwc_pad(x, out, NB, Cin, HW, CP, scale_dev=None):
    sc = 1.0 if scale_dev is None else float(scale_dev.reshape(-1)[0])
    o = out.view(NB, HW, CP)
    o.zero_()
    o[:, :, :Cin] = (x.float().reshape(NB, Cin, HW) * sc).permute(0, 2, 1)
    return out


def rows_to_nchw(x, ld, out, NB, Cout, HW):
    out.copy_(x.reshape(NB, HW, -1)[:, :, :Cout].permute(0, 2, 1).reshape(out.shape))
    return out


def add_inplace(y, x):
    y.add_(x)
    return y


def silu_inplace(y):
    y.copy_(F.silu(y.float()))
    return y


def timestep_embed(t_dev, t_stride, rows, dim, out, ld, col0=0):
    half = dim // 2
    t = t_dev.float().reshape(-1)
    tv = torch.stack([t[r * t_stride] for r in range(rows)])
    e = tv[:, None] * torch.exp(-math.log(10000.0) * torch.arange(half).float() / half)[None]
    emb = torch.cat([torch.cos(e), torch.sin(e)], dim=-1).to(out.dtype)
    flat = out.view(-1)                                   # the kernel addresses out[r * ld + col0 + k] in the flat buffer
    for r in range(rows):
        flat[r * ld + col0:r * ld + col0 + dim] = emb[r]
    return out


def skinny_linear(x, w, bias, out, M, N, K, silu_in=False, accumulate=False, act_in=None):
    act = {None: "silu" if silu_in else "none"}.get(act_in, act_in)
    xf = x.float()[:M, :K]
    xf = F.silu(xf) if act == "silu" else F.gelu(xf) if act == "gelu" else xf
    y = xf.to(x.dtype).float() @ w.float().T + (0 if bias is None else bias.float())
    out[:M, :N] = (out[:M, :N].float() + y) if accumulate else y
    return out


def softmax_rows(x, rows, cols):
    x[:rows, :cols] = torch.softmax(x[:rows, :cols].float(), dim=-1)
    return x


def perceiver_attn(q, kv, out, B, L, n_kv, heads, dim_head=64):
    inner = heads * dim_head
    s = dim_head ** -0.25
    sp = lambda t, n: t.float().reshape(B, n, heads, dim_head).transpose(1, 2)
    w = torch.softmax((sp(q, L) * s) @ (sp(kv[:, :inner], n_kv) * s).transpose(-1, -2), dim=-1)
    out.copy_((w @ sp(kv[:, inner:], n_kv)).transpose(1, 2).reshape(B * L, inner))
    return out


def cfg_sched_step(eps, ld_eps, x, x0_prev, x16, next_in, CP, B, HW, guidance, coef_table, step_dev):
    cx, ce, cp, kx, ke, sc = (float(v) for v in coef_table[int(step_dev.reshape(-1)[0])][:6])
    e = eps.float()[:, :4].reshape(2, B, HW, 4).permute(0, 1, 3, 2)                  # [2, B, 4, HW]
    e = e[0] + guidance * (e[1] - e[0])
    xv = x.reshape(B, 4, HW).clone()
    x0 = kx * xv + ke * e
    xn = cx * xv + ce * e + cp * x0_prev.reshape(B, 4, HW)
    x0_prev.copy_(x0.reshape(x0_prev.shape)); x.copy_(xn.reshape(x.shape)); x16.copy_(xn.reshape(x16.shape))
    if next_in is not None:
        rows = next_in.view(2, B, HW, CP)
        rows[:, :, :, :8] = 0
        rows[:, :, :, :4] = (xn * sc).permute(0, 2, 1)[None]


def latents_to_input(x, next_in, CP, B, HW, coef_table, step_dev=None, nsteps=1, keep_ch4_up=False):
    st = 0 if step_dev is None else min(int(step_dev.reshape(-1)[0]), nsteps - 1)
    rows = next_in.view(2, B, HW, CP)
    if not keep_ch4_up:
        rows[:, :, :, :8] = 0
    rows[:, :, :, :4] = (x.reshape(B, 4, HW) * float(coef_table[st][6])).permute(0, 2, 1)[None]


def advance_step(step_dev, t_dev, ts_table, n):
    s = int(step_dev.reshape(-1)[0]) + 1
    step_dev.fill_(s)
    t_dev.fill_(float(ts_table.reshape(-1)[min(s, n - 1)]))


def inpaint_blend(x, x16, image_latents, noise, mask, B, HW, blend_table, step_dev):
    ca, cn = (float(v) for v in blend_table.reshape(-1, 2)[int(step_dev.reshape(-1)[0])])
    m = mask.float().reshape(B, 1, HW)
    v = (1 - m) * (ca * image_latents.float().reshape(B, 4, HW) + cn * noise.float().reshape(B, 4, HW)) + m * x.reshape(B, 4, HW)
    x.copy_(v.reshape(x.shape)); x16.copy_(v.reshape(x16.shape))


def ensure_workspace(device=None):
    return None


_NAMES = ("gemm conv3x3 attn_self pack_cross_kv attn_cross gn_stats gn_apply gn_apply_ch gn_small gn_small_ok layernorm layernorm_rows upsample2x phase_split nchw_to_nhwc_pad "
          "rows_to_nchw add_inplace silu_inplace timestep_embed skinny_linear softmax_rows perceiver_attn cfg_sched_step latents_to_input "
          "advance_step inpaint_blend ensure_workspace").split()


def install(monkeypatch):
    """Route ``consistentid_b200.ops`` through the emulations and let CPU tensors pass the engines' ``is_cuda`` checks (test scope only)."""
    from consistentid_b200 import ops
    for n in _NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


def install_permanent():
    """Same as ``install`` for a throw-away process (spawned distributed workers): no restore."""
    from consistentid_b200 import ops
    for n in _NAMES:
        setattr(ops, n, globals()[n])
    torch.Tensor.is_cuda = property(lambda self: True)
