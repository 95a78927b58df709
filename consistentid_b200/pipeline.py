"""The ConsistentID denoising loop on the B200 engine.

Same control flow as the reference loop bodies
(pipline_StableDiffusion_ConsistentID.py:533-579, pipline_StableDiffusionXL_ConsistentID.py:608-667): per step
``cat([latents]*2)`` -> ``scale_model_input`` -> prompt switch at ``i <= start_merge_step`` -> UNet -> CFG combine ->
``scheduler.step``.  Here one step is ONE CUDA-graph replay: the UNet launch program followed by the fused
CFG + scheduler-step kernel, which also writes the next step's scaled, batch-duplicated NHWC UNet input, and a
1-thread bookkeeping kernel that advances the device-side step index / timestep.  No host sync inside the loop.

Batch semantics: B independent latents sharing one identity's prompt embeddings
== B batch-1 runs of the reference.
"""
from __future__ import annotations

import torch

from . import ops
from .scheduler import B200Scheduler
from .unet import B200UNet, CIN_PAD


class B200Denoiser:
    def __init__(self, unet: B200UNet, scheduler: B200Scheduler, use_cuda_graph=True):
        self.unet, self.scheduler, self.use_cuda_graph = unet, scheduler, use_cuda_graph
        self._graphs = {}
        self._graph_sig = None
        self.launches_per_step = None

    # ------------------------------------------------------------------ helpers
    def _dev16(self, t):
        return t.to(device=self.unet.device, dtype=self.unet.dtype, non_blocking=True)

    def _pair(self, neg, pos, B):
        """[2B, L, cad]: uncond rows first, then cond rows (``torch.cat([null, cond])``, :542-549)."""
        neg, pos = self._dev16(neg), self._dev16(pos)
        return torch.cat([neg.expand(B, *neg.shape[1:]), pos.expand(B, *pos.shape[1:])], dim=0).contiguous()

    def _step_eager(self, key, st):
        u = self.unet
        B, HW = st["B"], st["HW"]
        eps = u.forward(key)
        ops.cfg_sched_step(eps, 4, st["x"], st["x0"], st["x16"], u._buf("x_in", (2 * B * HW, CIN_PAD)), CIN_PAD, B, HW,
                           st["guidance"], st["coef"], st["step"])
        ops.advance_step(st["step"], u._buf("t_dev", (1,), torch.float32), st["ts"], st["n"])

    def _step(self, phase, key, st):
        if not self.use_cuda_graph:
            return self._step_eager(key, st)
        g = self._graphs.get(phase)
        if g is None:
            # warm-up once eagerly on a side stream (allocates every buffer, sets kernel attributes), restore state, capture
            snap = {k: st[k].clone() for k in ("x", "x0", "step")}
            t_dev = self.unet._buf("t_dev", (1,), torch.float32)
            x_in = self.unet._buf("x_in", (2 * st["B"] * st["HW"], CIN_PAD))
            t_snap, in_snap = t_dev.clone(), x_in.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._step_eager(key, st)
            torch.cuda.current_stream().wait_stream(s)
            for k, v in snap.items():
                st[k].copy_(v)
            t_dev.copy_(t_snap); x_in.copy_(in_snap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager(key, st)
            for k, v in snap.items():       # capture does not execute, but keep state explicit
                st[k].copy_(v)
            t_dev.copy_(t_snap); x_in.copy_(in_snap)
            self._graphs[phase] = g
        g.replay()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def __call__(self, latents, null_embeds, augmented_embeds, text_embeds, num_inference_steps=30, guidance_scale=5.0,
                 start_merge_step=0, neg_pooled=None, pooled_text_only=None, pooled_facial=None, add_time_ids=None,
                 null_embeds_facial=None, output_device=None, profile=False):
        """latents [B,4,h,w] already multiplied by ``scheduler.init_noise_sigma`` (host or device, any float dtype).
        SD1.5: (null, augmented, text_only) each [1,81,cad] (chunk(3) of prompt_embeds, :527-531).
        SDXL: additionally pooled embeds [1,1280] x3 and add_time_ids [1,6]; ``null_embeds_facial`` is the uncond prompt of
        the facial phase (pipline_StableDiffusionXL_ConsistentID.py:578-587) and defaults to ``null_embeds``.
        Returns final latents [B,4,h,w] in the engine dtype (on ``output_device`` if given)."""
        u, sch = self.unet, self.scheduler
        dev = u.device
        B, _, h, w = latents.shape
        HW, NB = h * w, 2 * B
        n = num_inference_steps
        sdxl = u.spec.addition_embed_type == "text_time"
        sig = (B, h, w, n, float(guidance_scale), sch.kind, u.ip_scale)
        if sig != self._graph_sig:
            self._graphs.clear()
            self._graph_sig = sig
        u.plan(NB, h, w)
        sch.set_timesteps(n, device=dev)
        coef, ts = sch.device_tables(dev)
        # ---- prompt phases (K/V caches + SDXL added-cond embedding), computed once per call
        null_f = null_embeds if null_embeds_facial is None else null_embeds_facial
        phases = {}
        need_text = start_merge_step >= 0
        need_aug = start_merge_step < n - 1
        def added(pos_pooled):
            if not sdxl:
                return None
            te = torch.cat([self._dev16(neg_pooled).expand(B, -1), self._dev16(pos_pooled).expand(B, -1)], 0).contiguous()
            ti = torch.cat([add_time_ids.to(dev, torch.float32).expand(B, -1)] * 2, 0).contiguous()
            return {"text_embeds": te, "time_ids": ti}
        if need_text:
            phases["text"] = u.set_prompt(self._pair(null_embeds, text_embeds, B), added(pooled_text_only), key="phase:text")
        if need_aug:
            phases["aug"] = u.set_prompt(self._pair(null_f, augmented_embeds, B), added(pooled_facial), key="phase:aug")
        # ---- per-run state
        st = {"B": B, "HW": HW, "n": n, "guidance": float(guidance_scale), "coef": coef, "ts": ts,
              "x": u._buf("lat32", (B, 4, HW), torch.float32), "x0": u._buf("lat_x0", (B, 4, HW), torch.float32),
              "x16": u._buf("lat16", (B, 4, HW)), "step": u._buf("step_dev", (1,), torch.int32)}
        st["x"].copy_(latents.reshape(B, 4, HW).to(dev, non_blocking=True))
        st["x0"].zero_()
        st["step"].zero_()
        u._buf("t_dev", (1,), torch.float32).copy_(ts[:1])
        ops.latents_to_input(st["x"], u._buf("x_in", (NB * HW, CIN_PAD)), CIN_PAD, B, HW, coef)
        if profile:                       # bench.py roofline pass: event-bracket every tensor-core launch of the step loop only
            ops.profile_begin()
        for i in range(n):
            phase = "text" if i <= start_merge_step else "aug"
            self._step(phase, phases[phase], st)
        if profile:
            self.last_profile = ops.profile_end()
        out = st["x16"].reshape(B, 4, h, w)
        if output_device is not None:
            return out.to(output_device)
        return out.clone()
