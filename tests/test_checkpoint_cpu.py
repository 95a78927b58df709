"""Checkpoint formats: ConsistentID-v1.bin sections, positional adapter keys, training-checkpoint split,
strict key/shape checking, spec inference from diffusers UNet weights.  Host logic only (no GPU)."""
import os

import pytest
import torch

from consistentid_b200 import checkpoint as ck
from consistentid_b200.arch import UNetSpec, attn_processor_names, param_shapes, sd15_spec, sdxl_spec

TINY = UNetSpec(block_out_channels=(64, 128, 256, 256), num_attention_heads=(2, 2, 4, 4), cross_attention_dim=128, sample_size=32, name="tiny_sd15")


def _adapter(spec, rank=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(s, generator=g) for k, s in param_shapes(spec, rank)[1].items()}


def test_bin_roundtrip_and_alias(tmp_path):
    ad = _adapter(TINY)
    extra = {"proj.weight": torch.randn(3, 3)}
    # loader spelling (pipline_StableDiffusion_ConsistentID.py:141-142) and converter spelling (evaluation/convert_weights.py:25)
    for key in ("image_proj", "image_proj_model"):
        p = tmp_path / f"ConsistentID-{key}.bin"
        torch.save({key: extra, "adapter_modules": ad, "FacialEncoder": {"w": torch.ones(2)}}, p)
        sec = ck.load_consistentid_checkpoint(p)
        assert set(sec) >= set(ck.SECTIONS)
        assert torch.equal(sec["image_proj"]["proj.weight"], extra["proj.weight"])
        assert all(torch.equal(sec["adapter_modules"][k], v) for k, v in ad.items())
        assert ck.check_adapter_modules(TINY, sec["adapter_modules"]) == 16


def test_training_checkpoint_split_and_safetensors(tmp_path):
    ad = _adapter(TINY, rank=8)
    flat = {f"adapter_modules.{k}": v for k, v in ad.items()}
    flat["unet.conv_in.weight"] = torch.zeros(1)             # frozen UNet entries are dropped
    flat["image_proj_model.norm.weight"] = torch.ones(4)
    flat["FacialEncoder.mlp.fc1.weight"] = torch.ones(4, 4)
    sec = ck.load_consistentid_checkpoint(flat)
    assert set(sec["adapter_modules"]) == set(ad) and "norm.weight" in sec["image_proj"] and "mlp.fc1.weight" in sec["FacialEncoder"]
    from safetensors.torch import save_file
    p = tmp_path / "ckpt.safetensors"
    save_file({k: v.contiguous() for k, v in flat.items()}, str(p))
    sec2 = ck.load_consistentid_checkpoint(p)
    assert all(torch.equal(sec2["adapter_modules"][k], v) for k, v in ad.items())
    assert ck.infer_lora_rank(sec2["adapter_modules"]) == 8


def test_strict_errors():
    ad = _adapter(TINY)
    with pytest.raises(KeyError):
        ck.load_consistentid_checkpoint({"image_proj": {}, "FacialEncoder": {}})
    miss = dict(ad); miss.pop("0.to_q_lora.down.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        ck.check_adapter_modules(TINY, miss, rank=16)
    extra = dict(ad); extra["999.to_k_ip.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        ck.check_adapter_modules(TINY, extra)
    bad = dict(ad); bad["1.to_k_ip.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        ck.check_adapter_modules(TINY, bad)
    # an SD1.5 adapter section cannot load into SDXL (different processor count / widths)
    with pytest.raises(RuntimeError):
        ck.check_adapter_modules(sdxl_spec(), {k: torch.empty(s, device="meta") for k, s in param_shapes(sd15_spec(), 128)[1].items()})


def test_positional_keys_follow_attn_processor_order():
    ad = _adapter(TINY)
    by_name = ck.adapter_modules_by_name(TINY, ad)
    names = attn_processor_names(TINY)
    assert list(by_name) == names
    for i, n in enumerate(names):
        keys = set(by_name[n])
        lora = {f"{p}_lora.{d}.weight" for p in ("to_q", "to_k", "to_v", "to_out") for d in ("down", "up")}
        assert keys == (lora | {"to_k_ip.weight", "to_v_ip.weight"} if n.endswith("attn2.processor") else lora), n
        assert by_name[n]["to_q_lora.down.weight"] is ad[f"{i}.to_q_lora.down.weight"]


def test_positional_keys_match_reference_processor_modules():
    """The positional section is exactly ModuleList(unet.attn_processors.values()).state_dict() of the (oracle-restated)
    UNet carrying the reference-named processors."""
    from oracle.synth import build_ref_unet
    from oracle.unet_ref import tiny_config
    cfg = tiny_config()
    unet = build_ref_unet(cfg, rank=16)
    spec = UNetSpec.from_config(cfg)
    ref_sd = torch.nn.ModuleList(unet.attn_processors.values()).state_dict()
    ours = ck.adapter_modules_from_processors(spec, unet.attn_processors)
    assert list(ours) == list(ref_sd)
    assert ck.check_adapter_modules(spec, ref_sd) == 16


def test_infer_spec_from_unet_weights(tmp_path):
    meta = lambda spec: {k: torch.empty(s, device="meta") for k, s in param_shapes(spec)[0].items()}
    assert ck.infer_spec(meta(sd15_spec())).in_channels == 4
    assert ck.infer_spec(meta(sd15_spec(in_channels=9))).in_channels == 9
    assert ck.infer_spec(meta(sdxl_spec())).addition_embed_type == "text_time"
    broken = meta(sd15_spec()); broken.pop("mid_block.resnets.0.conv1.weight")
    with pytest.raises(RuntimeError):
        ck.infer_spec(broken)
    # directory layouts of a diffusers model
    small = {"conv_in.weight": torch.zeros(2, 4, 3, 3)}
    os.makedirs(tmp_path / "model" / "unet")
    torch.save(small, tmp_path / "model" / "unet" / "diffusion_pytorch_model.bin")
    assert torch.equal(ck.load_unet_state_dict(tmp_path / "model")["conv_in.weight"], small["conv_in.weight"])
    assert torch.equal(ck.load_unet_state_dict(tmp_path / "model" / "unet")["conv_in.weight"], small["conv_in.weight"])
    with pytest.raises(FileNotFoundError):
        ck.load_unet_state_dict(tmp_path)
