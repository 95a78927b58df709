"""GPU: an engine built from checkpoint FILES (diffusers UNet directory + ConsistentID-v1.bin layout) equals the engine built from
the in-memory state dicts, and matches the oracle."""
import os

import pytest
import torch

from oracle import synth
from oracle.unet_ref import tiny_config
from tests.test_unet_gpu import _cmp, _engine_from_oracle


@pytest.mark.gpu
def test_engine_from_checkpoint_files(tmp_path):
    from consistentid_b200 import checkpoint as ck
    from consistentid_b200.arch import UNetSpec
    dtype = torch.float16
    cfg = tiny_config("sd15")
    ref = synth.build_ref_unet(cfg, rank=16)
    os.makedirs(tmp_path / "sd" / "unet")
    torch.save({k: v for k, v in ref.state_dict().items() if ".processor." not in k}, tmp_path / "sd" / "unet" / "diffusion_pytorch_model.bin")
    torch.save({"image_proj_model": {}, "FacialEncoder": {}, "adapter_modules": torch.nn.ModuleList(ref.attn_processors.values()).state_dict()},
               tmp_path / "ConsistentID-v1.bin")
    eng, sections = ck.build_engine(tmp_path / "sd", tmp_path / "ConsistentID-v1.bin", dtype=dtype, spec=UNetSpec.from_config(cfg))
    eng0 = _engine_from_oracle(ref, dtype, 16)
    # same packed, LoRA- and LayerNorm-folded arena bit for bit (compared as integers: the fp32 column sums live in the 16-bit arena as raw bits)
    assert torch.equal(eng.params.arena.view(torch.int16), eng0.params.arena.view(torch.int16))
    B, h = 1, cfg.sample_size
    null, aug, _ = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(2 * B, h, h, seed=5)
    ehs = torch.cat([null, aug])
    t = torch.tensor(401)
    with torch.no_grad():
        truth = ref(x, t, ehs).sample
        ref16 = synth.build_ref_unet(cfg, rank=16, dtype=dtype).cuda()
        for p in ref16.attn_processors.values():
            p.cuda()
        eager = ref16(x.cuda().to(dtype), t.cuda(), ehs.cuda().to(dtype)).sample
    out = eng(x.cuda().to(dtype), t, ehs.cuda().to(dtype), cross_attention_kwargs={}).sample
    torch.cuda.synchronize()
    _cmp("unet from checkpoint files", out, truth, eager)
