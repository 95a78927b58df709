"""B200UNet - the whole-step engine behind the reference's ``unet(...)`` call surface.

Drop-in for what the ConsistentID pipelines touch on ``pipe.unet``:
``unet(sample, t, encoder_hidden_states=, cross_attention_kwargs=, [added_cond_kwargs=],
[down_block_additional_residuals=, mid_block_additional_residual=]).sample``
(pipline_StableDiffusion_ConsistentID.py:552-557, pipline_StableDiffusionXL_ConsistentID.py:634-641),
``unet.config``, ``unet.in_channels``, ``unet.dtype``, ``unet.device``, ``unet.attn_processors``,
``unet.set_attn_processor`` (pipline_StableDiffusion_ConsistentID.py:152-174).

Everything numeric is a libcidb200 launch: NHWC 16-bit activations, LoRA folded into the base projections at load
(attention.py:138-146,236-250,282), fused QKV / GEGLU / residual / time-embedding epilogues, cross-attention K/V
(which depend only on the prompt) cached per prompt.  The launch sequence for a fixed (batch, H, W) is a static
"program" that can be replayed inside a CUDA graph (no host sync, no allocation).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import lib, ops
from .arch import UNetSpec, attn_processor_names, param_shapes, walk
from .lib import EPI_GEGLU, EPI_QKV
from .weights import TensorIdent, fold_layernorm, fold_lora, interleave_geglu, pack_conv3x3, same_tensors

N_TEXT_MAX, KROWS = 80, 96
CIN_PAD = 64


class _Params:
    """Packs every weight the engine needs, in the layout its kernels want, into ONE flat 16-bit arena
    (a single contiguous tensor: one NCCL broadcast moves the whole model, see dist.py)."""

    def __init__(self, spec: UNetSpec, unet_sd, adapter_sd, dtype, device, rank, lora_scale=1.0, kinds=("down", "mid", "up"),
                 finalize=True):
        """``adapter_sd`` None = no ConsistentID adapters (ControlNet keeps diffusers' default processors: no LoRA, no id branch);
        ``kinds`` selects the block groups that exist (ControlNet = down + mid, no head)."""
        self.spec, self.dtype, self.device = spec, dtype, device
        self._items = {}     # name -> (offset, shape)
        self._staged = []    # (name, tensor) before arena allocation
        self._f32 = set()    # names stored as raw fp32 bits
        self.kinds = kinds
        ushapes, ashapes = param_shapes(spec, rank)
        is_unet = "up" in kinds
        if is_unet:
            for n, s in ushapes.items():
                if n not in unet_sd or tuple(unet_sd[n].shape) != s:
                    raise KeyError(f"unet state_dict: missing/mis-shaped '{n}' (want {s})")
        if adapter_sd is not None:
            for n, s in ashapes.items():
                if n not in adapter_sd or tuple(adapter_sd[n].shape) != s:
                    raise KeyError(f"adapter_modules state_dict: missing/mis-shaped '{n}' (want {s}); keys are positional "
                                   f"'{{i}}.to_q_lora.down.weight' in unet.attn_processors order")
        U = lambda n: unet_sd[n].to(device=device, dtype=dtype)
        A = lambda n: adapter_sd[n].to(device=device, dtype=dtype)
        put = self._put
        T = spec.time_embed_dim
        # stem / head
        put("conv_in.w", pack_conv3x3(U("conv_in.weight"), CIN_PAD)); put("conv_in.b", U("conv_in.bias"))
        if is_unet:
            put("conv_out.w", pack_conv3x3(U("conv_out.weight"))); put("conv_out.b", U("conv_out.bias"))
            put("norm_out.g", U("conv_norm_out.weight")); put("norm_out.b", U("conv_norm_out.bias"))
        for k in ("time_embedding.linear_1", "time_embedding.linear_2"):
            put(k + ".w", U(k + ".weight")); put(k + ".b", U(k + ".bias"))
        if spec.addition_embed_type == "text_time" and is_unet:
            w1 = U("add_embedding.linear_1.weight")
            n_text = w1.shape[1] - 6 * spec.addition_time_embed_dim
            put("add1.w_text", w1[:, :n_text].contiguous()); put("add1.w_time", w1[:, n_text:].contiguous())
            put("add1.b", U("add_embedding.linear_1.bias"))
            put("add2.w", U("add_embedding.linear_2.weight")); put("add2.b", U("add_embedding.linear_2.bias"))
        pos_of = {n: i for i, n in enumerate(attn_processor_names(spec))}
        temb_w, temb_b, self.temb_off = [], [], {}
        off = 0
        for kind, i, layers, has_sampler in walk(spec):
            if kind not in kinds:
                continue
            for r, tf in layers:
                n = r.name
                put(n + ".n1.g", U(n + ".norm1.weight")); put(n + ".n1.b", U(n + ".norm1.bias"))
                put(n + ".c1.w", pack_conv3x3(U(n + ".conv1.weight"))); put(n + ".c1.b", U(n + ".conv1.bias"))
                put(n + ".n2.g", U(n + ".norm2.weight")); put(n + ".n2.b", U(n + ".norm2.bias"))
                put(n + ".c2.w", pack_conv3x3(U(n + ".conv2.weight"))); put(n + ".c2.b", U(n + ".conv2.bias"))
                if r.cin != r.cout:
                    put(n + ".sc.w", U(n + ".conv_shortcut.weight").reshape(r.cout, r.cin).contiguous())
                    put(n + ".sc.b", U(n + ".conv_shortcut.bias"))
                temb_w.append(U(n + ".time_emb_proj.weight")); temb_b.append(U(n + ".time_emb_proj.bias"))
                self.temb_off[n] = off; off += r.cout
                if tf is None:
                    continue
                t, C = tf.name, tf.channels
                put(t + ".norm.g", U(t + ".norm.weight")); put(t + ".norm.b", U(t + ".norm.bias"))
                put(t + ".pi.w", U(t + ".proj_in.weight").reshape(C, C).contiguous()); put(t + ".pi.b", U(t + ".proj_in.bias"))
                put(t + ".po.w", U(t + ".proj_out.weight").reshape(C, C).contiguous()); put(t + ".po.b", U(t + ".proj_out.bias"))
                for k in range(tf.layers):
                    b = f"{t}.transformer_blocks.{k}"
                    p1, p2 = pos_of[f"{b}.attn1.processor"], pos_of[f"{b}.attn2.processor"]

                    def folded(attn, pos, proj, out_name=None):
                        w = U(f"{b}.{attn}.{out_name or proj}.weight")
                        if adapter_sd is None:
                            return w
                        return fold_lora(w, A(f"{pos}.{proj}_lora.down.weight"), A(f"{pos}.{proj}_lora.up.weight"), lora_scale)
                    ln = lambda j: (U(f"{b}.norm{j}.weight"), U(f"{b}.norm{j}.bias"))
                    # norm1 / norm2 / norm3 are folded into the GEMM that consumes them (weights.fold_layernorm): gamma scales the weight
                    # columns, beta becomes a bias, and the column sums feed the epilogue's mean correction (cid_gemm ln_stats / ln_colsum)
                    wq, bq, cq = fold_layernorm(torch.cat([folded("attn1", p1, "to_q"), folded("attn1", p1, "to_k"), folded("attn1", p1, "to_v")], 0), None, *ln(1))
                    put(f"{b}.a1.qkv.w", wq); put(f"{b}.a1.qkv.b", bq); self._put_f32(f"{b}.a1.qkv.cs", cq)
                    put(f"{b}.a1.o.w", folded("attn1", p1, "to_out", "to_out.0")); put(f"{b}.a1.o.b", U(f"{b}.attn1.to_out.0.bias"))
                    wq, bq, cq = fold_layernorm(folded("attn2", p2, "to_q"), None, *ln(2))
                    put(f"{b}.a2.q.w", wq); put(f"{b}.a2.q.b", bq); self._put_f32(f"{b}.a2.q.cs", cq)
                    put(f"{b}.a2.k.w", folded("attn2", p2, "to_k")); put(f"{b}.a2.v.w", folded("attn2", p2, "to_v"))
                    if adapter_sd is not None:
                        put(f"{b}.a2.kip.w", A(f"{p2}.to_k_ip.weight")); put(f"{b}.a2.vip.w", A(f"{p2}.to_v_ip.weight"))
                    put(f"{b}.a2.o.w", folded("attn2", p2, "to_out", "to_out.0")); put(f"{b}.a2.o.b", U(f"{b}.attn2.to_out.0.bias"))
                    w, bb, _ = fold_layernorm(U(f"{b}.ff.net.0.proj.weight"), U(f"{b}.ff.net.0.proj.bias"), *ln(3))
                    tile = lib.gemm_tile_n(w.shape[0], EPI_GEGLU)
                    if tile < 0:
                        raise ValueError(f"GEGLU width {w.shape[0]} unsupported")
                    wi, bi = interleave_geglu(w, bb, tile)
                    put(f"{b}.ff1.w", wi); put(f"{b}.ff1.b", bi); self._put_f32(f"{b}.ff1.cs", wi.float().sum(dim=1))     # (column sums in the interleaved row order)
                    put(f"{b}.ff2.w", U(f"{b}.ff.net.2.weight")); put(f"{b}.ff2.b", U(f"{b}.ff.net.2.bias"))
            if has_sampler:
                nm = f"down_blocks.{i}.downsamplers.0.conv" if kind == "down" else f"up_blocks.{i}.upsamplers.0.conv"
                put(nm + ".w", pack_conv3x3(U(nm + ".weight"))); put(nm + ".b", U(nm + ".bias"))
        put("temb_all.w", torch.cat(temb_w, 0)); put("temb_all.b", torch.cat(temb_b, 0))
        self.temb_total = off
        if finalize:
            self._finalize()

    def _put(self, name, t):
        self._staged.append((name, t.contiguous()))

    def _put_f32(self, name, t):
        """fp32 tensor kept bit-exact inside the 16-bit arena (two arena elements per value; one broadcast still moves everything)."""
        self._f32.add(name)
        self._staged.append((name, t.float().contiguous().view(self.dtype)))

    def _finalize(self):
        total, offs = 0, []
        for name, t in self._staged:
            offs.append(total)
            total += (t.numel() + 127) // 128 * 128        # 256-byte aligned slots
        self.arena = torch.zeros(total, dtype=self.dtype, device=self.device)      # (zeros: the alignment padding between slots is part of the broadcast / checksums)
        for (name, t), o in zip(self._staged, offs):
            self.arena[o:o + t.numel()].copy_(t.reshape(-1))
            self._items[name] = (o, tuple(t.shape))
        self._staged = None

    def __getitem__(self, name):
        o, shape = self._items[name]
        n = 1
        for s in shape:
            n *= s
        v = self.arena[o:o + n]
        return v.view(torch.float32) if name in self._f32 else v.view(shape)


class B200UNet:
    def __init__(self, spec, unet_state_dict, adapter_state_dict, dtype=torch.float16, device="cuda", rank=128,
                 num_tokens=4, ip_scale=1.0, lora_scale=1.0):
        if not isinstance(spec, UNetSpec):
            spec = UNetSpec.from_config(spec)
        self.spec, self.dtype, self.device = spec, dtype, torch.device(device)
        self.config = SimpleNamespace(**{k: getattr(spec, k) for k in spec.__dataclass_fields__},
                                      attention_head_dim=spec.num_attention_heads, time_embed_dim=spec.time_embed_dim)
        self.in_channels = spec.in_channels
        self.num_tokens, self.ip_scale = num_tokens, ip_scale
        self.params = _Params(spec, unet_state_dict, adapter_state_dict, dtype, self.device, rank, lora_scale)
        ops.ensure_workspace(self.device)        # split-K scratch handed to the library now, never under CUDA-graph capture
        self._bufs = {}
        self._plan = None           # (NB, H, W)
        self._kv = {}               # prompt key -> per-layer (k_cat, vt_cat)
        self._aug = {}              # prompt key -> SDXL aug_emb [NB, T]
        self._active_key = None
        self._slots, self._ident, self._auto_next = {}, {}, 0
        self._graphs = {}
        self._gn_k = 0
        self.plan_epoch = 0         # bumped whenever plan() drops the buffers: CUDA graphs captured against them are stale
        self._procs = {n: _EngineProcessor(self, n) for n in attn_processor_names(spec)}

    # ------------------------------------------------------------------ diffusers-facing surface
    @property
    def attn_processors(self):
        return dict(self._procs)

    def set_attn_processor(self, processors):
        """Accepts the reference's processors (``Consistent_AttProcessor`` / ``Consistent_IPAttProcessor`` objects whose
        LoRA / ip weights were already folded into this engine at construction): only ``scale`` is live state."""
        if isinstance(processors, dict):
            for n, p in processors.items():
                if hasattr(p, "scale"):
                    self.set_ip_scale(float(p.scale))

    def set_ip_scale(self, scale: float):
        if scale != self.ip_scale:
            self.ip_scale = scale
            self._graphs.clear()
            self._program = None

    def __call__(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None,
                 down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=True):
        NB, Cin, H, W = sample.shape
        assert Cin == self.spec.in_channels and sample.dtype == self.dtype and sample.is_cuda
        self.plan(NB, H, W)
        key = self.set_prompt(encoder_hidden_states, added_cond_kwargs)
        t = torch.as_tensor(timestep, device=self.device).to(torch.float32).reshape(-1)[:1]
        self._buf("t_dev", (1,), torch.float32).copy_(t)
        ops.nchw_to_nhwc_pad(sample.contiguous(), self._buf("x_in", (NB * H * W, CIN_PAD)), NB, Cin, H * W, CIN_PAD)
        self.forward(key, residuals=(down_block_additional_residuals, mid_block_additional_residual))
        out = torch.empty((NB, self.spec.out_channels, H, W), dtype=self.dtype, device=self.device)
        ops.rows_to_nchw(self._buf("eps", (NB * H * W, 4)), 4, out, NB, self.spec.out_channels, H * W)
        return SimpleNamespace(sample=out) if return_dict else (out,)

    # ------------------------------------------------------------------ buffers / planning
    def _buf(self, name, shape, dtype=None, zero=False):
        key = (name, tuple(shape), dtype or self.dtype)
        b = self._bufs.get(key)
        if b is None:
            b = (torch.zeros if zero else torch.empty)(shape, dtype=dtype or self.dtype, device=self.device)
            self._bufs[key] = b
        return b

    def plan(self, NB, H, W):
        if self._plan != (NB, H, W):
            self._plan = (NB, H, W)
            self.plan_epoch = getattr(self, "plan_epoch", 0) + 1
            self._kv.clear(); self._aug.clear(); self._graphs.clear(); self._slots.clear(); self._ident.clear()
            self._bufs = {k: v for k, v in self._bufs.items() if k[0] in ("t_dev",)}
            self._buf("x_in", (NB * H * W, CIN_PAD), zero=True)
            self._buf("t_dev", (1,), torch.float32, zero=True)

    # ------------------------------------------------------------------ prompt-dependent state (once per prompt phase)
    def set_prompt(self, ehs, added_cond_kwargs=None, key=None):
        """Cache everything that depends only on the prompt: per attn2 layer K_cat / V_cat^T (to_k/to_v on the 77 text rows
        with LoRA folded, to_k_ip/to_v_ip on the id rows; attention.py:241-250,266-267) and, for SDXL, the
        ``add_embedding`` output.  Keyed by tensor identity so the two prompt phases of the delayed-conditioning switch
        (pipline_StableDiffusion_ConsistentID.py:542-549) are each computed once."""
        NB = self._plan[0]
        add = added_cond_kwargs or {}
        te, ti = add.get("text_embeds"), add.get("time_ids")
        if key is None:
            # drop-in path: the cache entry holds the keyed tensors themselves (weights.TensorIdent), so a later prompt can never alias
            # an earlier one through a recycled allocation.  A loop that re-creates its prompt tensor every step (the reference's does:
            # torch.cat at :542-549) therefore recomputes K/V every step - correct, ~5 small launches per attn2 layer; callers that
            # want the per-prompt cache pass an explicit ``key`` (B200Denoiser does).  Two rotating slots bound the memory.
            tensors = (ehs, te, ti)
            for k in ("auto:0", "auto:1"):
                if k in self._kv and same_tensors(self._ident.get(k), tensors):
                    self._active_key = k
                    return k
            key = "auto:%d" % self._auto_next
            self._auto_next ^= 1
            self._ident[key] = tuple(None if t is None else TensorIdent(t) for t in tensors)
        slot = self._slots.setdefault(key, len(self._slots))
        assert ehs.shape[0] == NB and ehs.dtype == self.dtype, (ehs.shape, NB, ehs.dtype)
        L, cad = ehs.shape[1], ehs.shape[2]
        n_ip = self.num_tokens
        n_text = L - n_ip
        if n_text > (N_TEXT_MAX if n_ip > 0 else KROWS):
            raise ValueError(f"{n_text} text tokens do not fit the {KROWS}-row key tile")
        text = ehs[:, :n_text].reshape(NB * n_text, cad).contiguous()
        ip = ehs[:, n_text:].reshape(NB * n_ip, cad).contiguous() if n_ip > 0 else None
        P = self.params
        kv = {}
        tmp = lambda nm, rows, C: self._buf(nm, (rows, C))
        for kind, i, layers, _ in walk(self.spec):
            if kind not in self.params.kinds:
                continue
            for _, tf in layers:
                if tf is None:
                    continue
                C, Hh = tf.channels, tf.heads
                for k in range(tf.layers):
                    b = f"{tf.name}.transformer_blocks.{k}"
                    kt, vt = tmp("kv_kt", NB * n_text, C), tmp("kv_vt", NB * n_text, C)
                    ops.gemm(text, P[f"{b}.a2.k.w"], kt); ops.gemm(text, P[f"{b}.a2.v.w"], vt)
                    ki = vi = None
                    if n_ip > 0:
                        ki, vi = tmp("kv_ki", NB * n_ip, C), tmp("kv_vi", NB * n_ip, C)
                        ops.gemm(ip, P[f"{b}.a2.kip.w"], ki); ops.gemm(ip, P[f"{b}.a2.vip.w"], vi)
                    k_cat = self._buf(f"kcat.{slot}.{b}", (NB, KROWS, C))
                    vt_cat = self._buf(f"vtcat.{slot}.{b}", (NB * Hh, C // Hh, KROWS))
                    ops.pack_cross_kv(kt, vt, ki, vi, k_cat, vt_cat, NB, C, Hh, n_text, n_ip)
                    kv[b] = (k_cat, vt_cat, n_text, n_ip)
        self._kv[key] = kv
        if self.spec.addition_embed_type == "text_time":
            T, D = self.spec.time_embed_dim, self.spec.addition_time_embed_dim
            assert te is not None and ti is not None, "SDXL UNet needs added_cond_kwargs text_embeds/time_ids"
            tid = ti.to(torch.float32).reshape(-1).contiguous()
            temb = self._buf("add_time", (NB, 6 * D))
            ops.timestep_embed(tid, 1, NB * 6, D, temb, D)
            h1 = self._buf("add_h1", (NB, T))
            te16 = te.to(self.dtype).contiguous()
            ops.skinny_linear(te16, P["add1.w_text"], P["add1.b"], h1, NB, T, te16.shape[1])
            ops.skinny_linear(temb, P["add1.w_time"], None, h1, NB, T, 6 * D, accumulate=True)
            aug = self._buf(f"aug.{slot}", (NB, T))
            ops.skinny_linear(h1, P["add2.w"], P["add2.b"], aug, NB, T, T, silu_in=True)
            self._aug[key] = aug
        self._active_key = key
        return key

    # ------------------------------------------------------------------ the per-step program
    # ------------------------------------------------------------------ building blocks of the launch program
    def _time_embedding(self, key):
        """SURVEY A.2 steps 1-2 -> (emb [NB,T], temb_all [NB, sum Cout]): all resnets' time_emb_proj in ONE skinny launch."""
        NB = self._plan[0]
        spec, P, buf = self.spec, self.params, self._buf
        T, c0 = spec.time_embed_dim, spec.block_out_channels[0]
        t_sin = buf("t_sin", (NB, c0))
        ops.timestep_embed(buf("t_dev", (1,), torch.float32), 0, NB, c0, t_sin, c0)
        e1 = buf("emb1", (NB, T))
        ops.skinny_linear(t_sin, P["time_embedding.linear_1.w"], P["time_embedding.linear_1.b"], e1, NB, T, c0)
        emb = buf("emb", (NB, T))
        if spec.addition_embed_type == "text_time":
            emb.copy_(self._aug[key])
            ops.skinny_linear(e1, P["time_embedding.linear_2.w"], P["time_embedding.linear_2.b"], emb, NB, T, T, silu_in=True, accumulate=True)
        else:
            ops.skinny_linear(e1, P["time_embedding.linear_2.w"], P["time_embedding.linear_2.b"], emb, NB, T, T, silu_in=True)
        temb_all = buf("temb_all", (NB, P.temb_total))
        ops.skinny_linear(emb, P["temb_all.w"], P["temb_all.b"], temb_all, NB, P.temb_total, T, silu_in=True)
        return emb, temb_all

    # ---- GroupNorm statistics fused into the producers' epilogues -------------------------------------------------------------
    STATS_ARENA_FLOATS = 4 << 20        # 16 MB: >= 2 * NB * sum(C) over every normalised tensor of a forward at the BASELINE batches

    def _stats_begin(self):
        """Start of a forward: one memset zeroes the per-(sample, channel) statistics slots the producers of this forward will fill."""
        arena = self._buf("gn_chan_stats", (self.STATS_ARENA_FLOATS,), torch.float32)
        used = getattr(self, "_stats_used", None)
        (arena if used is None else arena[:used]).zero_()
        self._stats_off, self._stats = 0, {}

    def _stats_slot(self, out, HW, C):
        """Statistics slot for tensor ``out`` [NB*HW, C] about to be produced by a conv / GEMM, or None when the fused path does not apply
        (a 128-row tile would span two samples: HW % 128 != 0 - the 8x8 level; those tensors take the standalone statistics kernel)."""
        NB = self._plan[0]
        n = NB * C * 2
        if HW % 128 != 0 or self._stats_off + n > self.STATS_ARENA_FLOATS:
            self._stats.pop(out.data_ptr(), None)
            return None
        v = self._buf("gn_chan_stats", (self.STATS_ARENA_FLOATS,), torch.float32)[self._stats_off:self._stats_off + n]
        self._stats_off += (n + 63) // 64 * 64
        self._stats[out.data_ptr()] = v
        return v

    def _groupnorm(self, x1, C1, x2, C2, HW, g, b, eps, silu, out):
        NB, G = self._plan[0], self.spec.norm_num_groups
        s1 = self._stats.get(x1.data_ptr())
        s2 = self._stats.get(x2.data_ptr()) if x2 is not None else None
        if s1 is not None and (x2 is None or s2 is not None):
            ops.gn_apply_ch(x1, C1, s1, x2, C2, s2, NB, HW, G, g, b, eps, silu, out)
            return
        if ops.gn_small_ok(C1, C2, HW, G):
            # small tensors whose statistics could not ride on the producer (the 8x8 level): statistics + apply in ONE launch, one CTA per
            # (sample, group) with its slab in registers - the stats + apply pair cost ~40 us per GroupNorm for 2.6 MB of data
            ops.gn_small(x1, C1, x2, C2, NB, HW, G, g, b, eps, silu, out)
            return
        # two alternating statistics buffers: the first GroupNorm of a forward zeroes its own (one memset), every apply zeroes the
        # buffer the next GroupNorm will accumulate into
        k = self._gn_k
        self._gn_k = k + 1
        sums, nxt = self._buf(f"gn_sums{k & 1}", (NB, G, 2), torch.float32), self._buf(f"gn_sums{(k + 1) & 1}", (NB, G, 2), torch.float32)
        ops.gn_stats(x1, C1, x2, C2, NB, HW, G, sums, zero_sums=(k == 0))
        ops.gn_apply(x1, C1, x2, C2, NB, HW, G, sums, g, b, eps, silu, out, zero_next=nxt)

    def _resnet(self, r, x, skip, h, w, out_name, temb_all):
        """ResnetBlock2D on NHWC rows; ``skip`` = second source of the virtual channel concat (up path)."""
        NB = self._plan[0]
        spec, P, buf = self.spec, self.params, self._buf
        HW = h * w
        M = NB * HW
        c_x = r.cin - r.skip_ch
        act = buf("act", (M, r.cin))
        self._groupnorm(x, c_x, skip, r.skip_ch, HW, P[r.name + ".n1.g"], P[r.name + ".n1.b"], spec.norm_eps, True, act)
        h1 = buf("res_h1", (M, r.cout))
        o = P.temb_off[r.name]
        ops.conv3x3(act, P[r.name + ".c1.w"], h1, NB, h, w, r.cin, r.cout, bias=P[r.name + ".c1.b"], rowbias=temb_all[:, o:o + r.cout],
                    chan_stats=self._stats_slot(h1, HW, r.cout))
        act2 = buf("act", (M, r.cout))
        self._groupnorm(h1, r.cout, None, 0, HW, P[r.name + ".n2.g"], P[r.name + ".n2.b"], spec.norm_eps, True, act2)
        if r.cin != r.cout:
            res = buf("res_sc", (M, r.cout))
            ops.gemm(x, P[r.name + ".sc.w"], res, bias=P[r.name + ".sc.b"], a2=skip)
        else:
            res = x
        out = buf(out_name, (M, r.cout))
        ops.conv3x3(act2, P[r.name + ".c2.w"], out, NB, h, w, r.cout, r.cout, bias=P[r.name + ".c2.b"], residual=res,
                    chan_stats=self._stats_slot(out, HW, r.cout))
        return out

    def _transformer(self, tf, x, h, w, out_name, kv):
        """Transformer2DModel: GN -> proj_in -> [LN, self-attn, LN, decoupled cross-attn, LN, GEGLU FF] x layers -> proj_out + x."""
        NB = self._plan[0]
        P, buf = self.params, self._buf
        HW = h * w
        M = NB * HW
        C, Hh = tf.channels, tf.heads
        d = C // Hh
        gn = buf("act", (M, C))
        self._groupnorm(x, C, None, 0, HW, P[tf.name + ".norm.g"], P[tf.name + ".norm.b"], 1e-6, False, gn)
        t = buf("tf_h", (M, C))
        # LayerNorm statistics travel from the GEMM that writes the residual stream (row_stats: per-row sum / sum of squares, accumulated in its
        # epilogue) to the GEMM that consumes LayerNorm(t) (ln=...: weights carry gamma / beta, the epilogue applies mean / rstd): no LayerNorm pass.
        n_ln = 3 * tf.layers
        ln_stats = buf(f"ln_stats.{n_ln}", (n_ln, M, 2), torch.float32)
        ln_stats.zero_()
        eps = 1e-5
        ops.gemm(gn, P[tf.name + ".pi.w"], t, bias=P[tf.name + ".pi.b"], row_stats=ln_stats[0])
        qk = buf("tf_qk", (M, 2 * C))
        vt, ao = buf("tf_vt", (NB * Hh, d, HW)), buf("tf_ao", (M, C))
        q, ffm = buf("tf_q", (M, C)), buf("tf_ffm", (M, 4 * C))
        for k in range(tf.layers):
            b = f"{tf.name}.transformer_blocks.{k}"
            s1, s2, s3 = ln_stats[3 * k], ln_stats[3 * k + 1], ln_stats[3 * k + 2]
            s_next = ln_stats[3 * k + 3] if k + 1 < tf.layers else None
            ops.gemm(t, P[b + ".a1.qkv.w"], qk, bias=P[b + ".a1.qkv.b"], epi=EPI_QKV, vt=vt, n_split=2 * C, heads=Hh, hdim=d, ntok=HW,
                     ln=(s1, P[b + ".a1.qkv.cs"], eps))
            ops.attn_self(qk[:, :C], qk[:, C:], vt, ao, NB, Hh, HW, d)
            ops.gemm(ao, P[b + ".a1.o.w"], t, bias=P[b + ".a1.o.b"], residual=t, row_stats=s2)
            ops.gemm(t, P[b + ".a2.q.w"], q, bias=P[b + ".a2.q.b"], ln=(s2, P[b + ".a2.q.cs"], eps))
            k_cat, vt_cat, n_text, n_ip = kv[b]
            ops.attn_cross(q, k_cat, vt_cat, ao, NB, Hh, HW, d, n_text, n_ip, self.ip_scale)
            ops.gemm(ao, P[b + ".a2.o.w"], t, bias=P[b + ".a2.o.b"], residual=t, row_stats=s3)
            ops.gemm(t, P[b + ".ff1.w"], ffm, bias=P[b + ".ff1.b"], epi=EPI_GEGLU, ln=(s3, P[b + ".ff1.cs"], eps))
            ops.gemm(ffm, P[b + ".ff2.w"], t, bias=P[b + ".ff2.b"], residual=t, row_stats=s_next)
        out = buf(out_name, (M, C))
        st = self._stats_slot(out, HW, C)
        ops.gemm(t, P[tf.name + ".po.w"], out, bias=P[tf.name + ".po.b"], residual=x, chan_stats=st, stats_rows=HW if st is not None else 0)
        return out

    def _downsample(self, x, i, h, w, c):
        NB, P, buf = self._plan[0], self.params, self._buf
        ps = buf("phase", (NB * h * w, c))
        ops.phase_split(x, ps, NB, h, w, c)
        nm = f"down_blocks.{i}.downsamplers.0.conv"
        out = buf(f"h.{nm}", (NB * (h // 2) * (w // 2), c))
        ops.conv3x3(ps, P[nm + ".w"], out, NB, h // 2, w // 2, c, c, bias=P[nm + ".b"], stride2=True,
                    chan_stats=self._stats_slot(out, (h // 2) * (w // 2), c))
        return out

    def _add_residual(self, dst, res, c):
        """dst[M, c] += ControlNet residual.  ``res`` is NCHW (drop-in API) or NHWC rows; a residual computed for B samples is
        applied to both CFG halves of a 2B batch (the reference broadcasts [1,...] onto [2,...],
        pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:389-425)."""
        rows = res if res.ndim == 2 else self._to_rows(res, c)
        M = dst.shape[0]
        if rows.shape[0] == M:
            ops.add_inplace(dst, rows)
        else:
            assert rows.shape[0] * 2 == M, (rows.shape, dst.shape)
            ops.add_inplace(dst[:M // 2], rows)
            ops.add_inplace(dst[M // 2:], rows)

    # ------------------------------------------------------------------ the per-step program
    def forward(self, key=None, residuals=(None, None)):
        """One UNet evaluation on the NHWC input already staged in buffer ``x_in`` (timestep in ``t_dev``); result rows
        in buffer ``eps`` [NB*H*W, 4]."""
        key = key if key is not None else self._active_key
        NB, H, W = self._plan
        spec, P, buf = self.spec, self.params, self._buf
        kv = self._kv[key]
        self._gn_k = 0
        self._stats_begin()
        down_res, mid_res = residuals
        emb, temb_all = self._time_embedding(key)
        # --- stem
        h, w = H, W
        c0 = spec.block_out_channels[0]
        x = buf("h.conv_in", (NB * h * w, c0))
        ops.conv3x3(buf("x_in", (NB * H * W, CIN_PAD)), P["conv_in.w"], x, NB, h, w, CIN_PAD, c0, bias=P["conv_in.b"],
                    chan_stats=self._stats_slot(x, h * w, c0))
        skips = [(x, c0)]
        for kind, i, layers, has_sampler in walk(spec):
            if kind == "mid" and down_res is not None:
                # ControlNet residuals go onto COPIES of the skips: the tensor flowing into the mid block stays unmodified
                # (diffusers 0.23 UNet2DConditionModel.forward; pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:418-425)
                for idx, ((sk, c), r_) in enumerate(zip(list(skips), down_res)):
                    cp = buf(f"skipres.{idx}", tuple(sk.shape))
                    cp.copy_(sk)
                    self._add_residual(cp, r_, c)
                    skips[idx] = (cp, c)
            for j, (r, tf) in enumerate(layers):
                skip = None
                if kind == "up":
                    skip, _ = skips.pop()
                x = self._resnet(r, x, skip, h, w, f"h.{r.name}", temb_all)
                if tf is not None:
                    x = self._transformer(tf, x, h, w, f"h.{tf.name}", kv)
                if kind == "down":
                    skips.append((x, r.cout))
            if kind == "mid" and mid_res is not None:
                xm = buf("midres", tuple(x.shape))
                xm.copy_(x)
                self._add_residual(xm, mid_res, layers[-1][0].cout)
                x = xm
            if has_sampler:
                c = layers[-1][0].cout
                if kind == "down":
                    x = self._downsample(x, i, h, w, c)
                    h, w = h // 2, w // 2
                    skips.append((x, c))
                else:
                    up = buf("upsampled", (NB * 4 * h * w, c))
                    ops.upsample2x(x, up, NB, h, w, c)
                    h, w = 2 * h, 2 * w
                    nm = f"up_blocks.{i}.upsamplers.0.conv"
                    x2 = buf(f"h.{nm}", (NB * h * w, c))
                    ops.conv3x3(up, P[nm + ".w"], x2, NB, h, w, c, c, bias=P[nm + ".b"], chan_stats=self._stats_slot(x2, h * w, c))
                    x = x2
        # --- head
        act = buf("act", (NB * h * w, c0))
        self._groupnorm(x, c0, None, 0, h * w, P["norm_out.g"], P["norm_out.b"], spec.norm_eps, True, act)
        eps = buf("eps", (NB * H * W, 4))
        ops.conv3x3(act, P["conv_out.w"], eps, NB, h, w, c0, spec.out_channels, bias=P["conv_out.b"])
        self._stats_used = self._stats_off
        return eps

    def _to_rows(self, nchw, c):
        """ControlNet residual given as NCHW tensor -> NHWC rows (host-side plumbing; config 5 only)."""
        n, cc, hh, ww = nchw.shape
        assert cc == c
        return nchw.permute(0, 2, 3, 1).reshape(n * hh * ww, c).contiguous()


class _EngineProcessor:
    """What ``unet.attn_processors`` returns for a B200UNet: a handle carrying the mutable ``scale`` attribute of the
    reference's Consistent_IPAttProcessor (``set_scale``, pipline_StableDiffusion_ConsistentID.py:211-214)."""

    def __init__(self, engine, name):
        self._engine, self.name = engine, name
        self.is_cross = name.endswith("attn2.processor")

    @property
    def scale(self):
        return self._engine.ip_scale

    @scale.setter
    def scale(self, v):
        if self.is_cross:
            self._engine.set_ip_scale(float(v))
