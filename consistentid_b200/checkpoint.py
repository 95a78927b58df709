"""Checkpoint formats of the reference -> the inputs of the B200 engine.

What the reference reads and writes:

* ``ConsistentID-v1.bin`` / ``ConsistentID_SDXL-v1.bin``: ``torch.save`` of a dict with three sections, consumed by
  ``load_ConsistentID_model`` (pipline_StableDiffusion_ConsistentID.py:134-144):
  ``state_dict["FacialEncoder"]``, ``state_dict["image_proj"]`` and ``state_dict["adapter_modules"]``.  The adapter section
  is loaded with ``ModuleList(unet.attn_processors.values()).load_state_dict(..., strict=True)``, so its keys are
  POSITIONAL: ``"{i}.to_q_lora.down.weight"`` ... with ``i`` the index of the processor in ``unet.attn_processors``
  (module registration order down_blocks -> up_blocks -> mid_block, attn1 before attn2; ``arch.attn_processor_names``).
* the training checkpoint ``pytorch_model.bin`` (flat, prefixes ``unet.`` / ``image_proj_model.`` / ``adapter_modules.`` /
  ``FacialEncoder.``) and the script that strips it, evaluation/convert_weights.py:14-25, which names the projector section
  ``image_proj_model`` (the loader asks for ``image_proj``): both spellings are accepted here.
* the diffusers UNet weights (``<model>/unet/diffusion_pytorch_model.{safetensors,bin}``), a flat state_dict whose names are
  ``arch.param_shapes(spec)[0]``.

The ``.safetensors`` branch of the reference loader (``id_encoder.`` / ``lora_weights.`` prefixes, :124-131) cannot feed the three
sections it then indexes; the flat safetensors layout accepted here is ``<section>.<key>`` with the section names above.

Everything here is host code on CPU tensors; the packing / LoRA folding into the weight arena happens in ``unet._Params``.
"""
from __future__ import annotations

import os
import re

import torch

from .arch import UNetSpec, attn_processor_names, param_shapes, sd15_spec, sdxl_spec

SECTIONS = ("adapter_modules", "image_proj", "FacialEncoder")
_ALIASES = {"image_proj_model": "image_proj"}


def _read_file(path):
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(str(path), device="cpu")
    return torch.load(str(path), map_location="cpu", weights_only=True)


def split_training_checkpoint(flat_sd):
    """Flat training state_dict -> the three sections (evaluation/convert_weights.py:14-25; the frozen ``unet.`` entries are dropped)."""
    out = {s: {} for s in SECTIONS}
    for k, v in flat_sd.items():
        head, _, rest = k.partition(".")
        head = _ALIASES.get(head, head)
        if head in out and rest:
            out[head][rest] = v
    return out


def load_consistentid_checkpoint(src):
    """``src``: path to a ``.bin`` / ``.safetensors`` file or an already loaded dict (the reference accepts both,
    pipline_StableDiffusion_ConsistentID.py:111-137).  Returns ``{"adapter_modules", "image_proj", "FacialEncoder"}``;
    a missing adapter section is an error (as ``state_dict["adapter_modules"]`` would raise KeyError), the two encoder
    sections may be absent when only the denoising path is used."""
    sd = src if isinstance(src, dict) else _read_file(src)
    if any(isinstance(v, dict) for v in sd.values()):
        out = {_ALIASES.get(k, k): v for k, v in sd.items() if isinstance(v, dict)}
    else:
        out = split_training_checkpoint(sd)
    if not out.get("adapter_modules"):
        raise KeyError("checkpoint has no 'adapter_modules' section")
    for s in SECTIONS:
        out.setdefault(s, {})
    return out


def infer_lora_rank(adapter_sd):
    for k, v in adapter_sd.items():
        if k.endswith("_lora.down.weight"):
            return int(v.shape[0])
    raise KeyError("adapter_modules holds no '*_lora.down.weight' entry")


def check_adapter_modules(spec: UNetSpec, adapter_sd, rank=None):
    """``load_state_dict(strict=True)`` semantics against the processors ``set_ip_adapter`` would create
    (pipline_StableDiffusion_ConsistentID.py:152-174): raises RuntimeError listing missing / unexpected keys and size
    mismatches; returns the LoRA rank."""
    rank = rank or infer_lora_rank(adapter_sd)
    want = param_shapes(spec, rank)[1]
    missing = [k for k in want if k not in adapter_sd]
    unexpected = [k for k in adapter_sd if k not in want]
    bad = [f"{k}: checkpoint {tuple(adapter_sd[k].shape)} vs model {want[k]}" for k in want
           if k in adapter_sd and tuple(adapter_sd[k].shape) != want[k]]
    if missing or unexpected or bad:
        msg = ["Error(s) in loading adapter_modules state_dict:"]
        if missing:
            msg.append(f"  Missing key(s): {missing[:8]}{' ...' if len(missing) > 8 else ''} ({len(missing)})")
        if unexpected:
            msg.append(f"  Unexpected key(s): {unexpected[:8]}{' ...' if len(unexpected) > 8 else ''} ({len(unexpected)})")
        if bad:
            msg.append(f"  size mismatch: {bad[:8]}{' ...' if len(bad) > 8 else ''} ({len(bad)})")
        raise RuntimeError("\n".join(msg))
    return rank


def adapter_modules_by_name(spec: UNetSpec, adapter_sd):
    """Positional section -> ``{processor name: {parameter name: tensor}}`` (parameter names as in attention.py:105-108,194-200)."""
    names = attn_processor_names(spec)
    out = {n: {} for n in names}
    for k, v in adapter_sd.items():
        m = re.match(r"(\d+)\.(.+)$", k)
        if m is None or int(m.group(1)) >= len(names):
            raise KeyError(f"adapter_modules key '{k}' is not '<position>.<parameter>' with position < {len(names)}")
        out[names[int(m.group(1))]][m.group(2)] = v
    return out


def adapter_modules_from_processors(spec: UNetSpec, processors):
    """``{name: nn.Module}`` (e.g. ``unet.attn_processors`` of a diffusers UNet carrying the reference processors) ->
    the positional section, i.e. what ``ModuleList(processors.values()).state_dict()`` yields in ``attn_processors`` order."""
    out = {}
    for i, n in enumerate(attn_processor_names(spec)):
        for k, v in processors[n].state_dict().items():
            out[f"{i}.{k}"] = v
    return out


def load_unet_state_dict(src):
    """diffusers UNet weights: a state_dict, a weights file, a ``unet/`` directory or a model directory containing ``unet/``."""
    if isinstance(src, dict):
        return src
    p = str(src)
    if os.path.isdir(p):
        for sub in ("", "unet"):
            for fn in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.bin"):
                f = os.path.join(p, sub, fn)
                if os.path.exists(f):
                    return _read_file(f)
        raise FileNotFoundError(f"no diffusion_pytorch_model.* under {p}")
    return _read_file(p)


def infer_spec(unet_sd) -> UNetSpec:
    """SD1.5 (4- or 9-channel inpaint) vs SDXL from the weights themselves, then a full name/shape check."""
    cin = int(unet_sd["conv_in.weight"].shape[1])
    spec = sdxl_spec() if "add_embedding.linear_1.weight" in unet_sd else sd15_spec(in_channels=cin)
    want = param_shapes(spec)[0]
    bad = [k for k, s in want.items() if k not in unet_sd or tuple(unet_sd[k].shape) != s]
    if bad:
        raise RuntimeError(f"UNet state_dict does not match {spec.name}: {bad[:6]}{' ...' if len(bad) > 6 else ''} ({len(bad)})")
    return spec


def build_engine(unet_src, checkpoint_src, dtype=torch.float16, device="cuda", spec=None, **kw):
    """UNet weights + ConsistentID checkpoint -> ``B200UNet`` with the adapters folded in (the engine-level counterpart of
    ``ConsistentIDPipeline.load_ConsistentID_model`` + ``set_ip_adapter``).  Returns ``(unet, sections)``; ``sections`` keeps
    the ``image_proj`` / ``FacialEncoder`` weights for the embedding producers."""
    from .unet import B200UNet
    usd = load_unet_state_dict(unet_src)
    spec = spec or infer_spec(usd)
    sections = load_consistentid_checkpoint(checkpoint_src)
    rank = check_adapter_modules(spec, sections["adapter_modules"])
    return B200UNet(spec, usd, sections["adapter_modules"], dtype=dtype, device=device, rank=rank, **kw), sections
