"""CPU oracle: the denoising-loop bodies of the reference pipelines, driven with
pre-computed prompt / ID embeddings (the preprocessing that produces them is out of scope).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  pipline_StableDiffusion_ConsistentID.py:533-579     SD1.5 loop (cat latents x2, scale_model_input,
        prompt switch at ``i <= start_merge_step``, unet, CFG combine, scheduler.step)
  pipline_StableDiffusionXL_ConsistentID.py:608-667   SDXL loop (+ add_text_embeds / add_time_ids switch)
Batch extension (SURVEY.md 8a "Batch note"): the reference runs batch 1 only; batch B here means
B independent latents sharing one identity's embeddings, i.e. exactly B batch-1 reference runs.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def denoise_sd15(unet, scheduler, latents, null_embeds, augmented_embeds, text_embeds, num_inference_steps,
                 guidance_scale=5.0, start_merge_step=0, callback=None):
    """latents [B,4,h,w] (already scaled by init_noise_sigma); *_embeds [1,81,cad]."""
    b = latents.shape[0]
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    for i, t in enumerate(scheduler.timesteps):
        x_in = torch.cat([latents] * 2)
        x_in = scheduler.scale_model_input(x_in, t)
        cond = text_embeds if i <= start_merge_step else augmented_embeds
        ehs = torch.cat([null_embeds.expand(b, -1, -1), cond.expand(b, -1, -1)], dim=0)
        eps = unet(x_in, t, encoder_hidden_states=ehs, cross_attention_kwargs={}).sample
        eps_u, eps_c = eps.chunk(2)
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        latents = scheduler.step(eps, t, latents).prev_sample
        if callback is not None:
            callback(i, t, latents)
    return latents


@torch.no_grad()
def denoise_sdxl(unet, scheduler, latents, neg_text_only, pos_text_only, neg_facial, pos_facial,
                 neg_pooled, pooled_text_only, pooled_facial, add_time_ids, num_inference_steps,
                 guidance_scale=7.5, start_merge_step=0, callback=None):
    """Embeds [1,81,2048]; pooled [1,1280]; add_time_ids [1,6] (same for both CFG halves)."""
    b = latents.shape[0]
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    time_ids = torch.cat([add_time_ids.expand(b, -1)] * 2, dim=0)
    for i, t in enumerate(scheduler.timesteps):
        x_in = torch.cat([latents] * 2)
        x_in = scheduler.scale_model_input(x_in, t)
        if i <= start_merge_step:
            ehs = torch.cat([neg_text_only.expand(b, -1, -1), pos_text_only.expand(b, -1, -1)], dim=0)
            pooled = torch.cat([neg_pooled.expand(b, -1), pooled_text_only.expand(b, -1)], dim=0)
        else:
            ehs = torch.cat([neg_facial.expand(b, -1, -1), pos_facial.expand(b, -1, -1)], dim=0)
            pooled = torch.cat([neg_pooled.expand(b, -1), pooled_facial.expand(b, -1)], dim=0)
        eps = unet(x_in, t, encoder_hidden_states=ehs, cross_attention_kwargs={},
                   added_cond_kwargs={"text_embeds": pooled, "time_ids": time_ids}).sample
        eps_u, eps_c = eps.chunk(2)
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        latents = scheduler.step(eps, t, latents).prev_sample
        if callback is not None:
            callback(i, t, latents)
    return latents


@torch.no_grad()
def denoise_controlnet_inpaint(unet, controlnet, scheduler, latents, null_embeds, augmented_embeds, text_embeds, control_image,
                               image_latents, noise, mask, num_inference_steps, guidance_scale=5.0, start_merge_step=0,
                               conditioning_scale=1.0, masked_image_latents=None):
    """pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:375-449 (strength 1.0): ControlNet on the cond half only,
    its residuals added to BOTH CFG halves of the UNet (the reference broadcasts batch 1 onto batch 2), CFG, scheduler step and,
    for the 4-channel UNet, the latent blend with the re-noised original.  ``mask`` [B,1,h,w]: 1 = repaint.
    9-channel UNet: pass ``masked_image_latents`` [B,4,h,w] (concatenated with the mask every step, :414-415)."""
    b = latents.shape[0]
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    ts = scheduler.timesteps
    for i, t in enumerate(ts):
        x_in = scheduler.scale_model_input(torch.cat([latents] * 2), t)
        cond = text_embeds if i <= start_merge_step else augmented_embeds
        ehs = torch.cat([null_embeds.expand(b, -1, -1), cond.expand(b, -1, -1)], dim=0)
        down2 = mid2 = None
        if controlnet is not None:     # None = the plain inpaint loop, pipelines/StableDIffusionInpaint_ConsistentID.py:305-359
            c_in = scheduler.scale_model_input(latents, t)
            down, mid = controlnet(c_in, t, encoder_hidden_states=cond.expand(b, -1, -1), controlnet_cond=control_image,
                                   conditioning_scale=conditioning_scale)
            down2 = [torch.cat([d, d]) for d in down]
            mid2 = torch.cat([mid, mid])
        if masked_image_latents is not None:
            x_in = torch.cat([x_in, torch.cat([mask] * 2), torch.cat([masked_image_latents] * 2)], dim=1)
        eps = unet(x_in, t, encoder_hidden_states=ehs, cross_attention_kwargs={}, down_block_additional_residuals=down2,
                   mid_block_additional_residual=mid2).sample
        eps_u, eps_c = eps.chunk(2)
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        latents = scheduler.step(eps, t, latents).prev_sample
        if masked_image_latents is None:
            proper = image_latents
            if i < len(ts) - 1:
                proper = scheduler.add_noise(image_latents, noise, torch.tensor([int(ts[i + 1])]))
            latents = (1 - mask) * proper + mask * latents
    return latents


def denoise_inpaint(unet, scheduler, latents, null_embeds, augmented_embeds, text_embeds, image_latents, noise, mask, num_inference_steps,
                    guidance_scale=5.0, start_merge_step=0, masked_image_latents=None):
    """pipelines/StableDIffusionInpaint_ConsistentID.py:305-359 (strength 1.0): the loop above without a ControlNet (9-channel concat :320-321,
    4-channel latent blend :340-352)."""
    return denoise_controlnet_inpaint(unet, None, scheduler, latents, null_embeds, augmented_embeds, text_embeds, None, image_latents, noise, mask,
                                      num_inference_steps, guidance_scale=guidance_scale, start_merge_step=start_merge_step,
                                      masked_image_latents=masked_image_latents)
