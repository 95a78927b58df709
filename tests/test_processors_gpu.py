"""GPU: the drop-in processors (consistentid_b200/processors.py) behind the diffusers AttnProcessor protocol reproduce the
reference processors' outputs: golden vectors generated from the verbatim reference (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle.unet_ref import Attention

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "processors_golden.pt")


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1, 2])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_processors_match_reference_golden(idx, dtype):
    from consistentid_b200.processors import Consistent_AttProcessor, Consistent_IPAttProcessor
    case = torch.load(GOLDEN)[idx]
    m = case["meta"]
    a1 = Attention(m["C"], None, m["heads"], m["C"] // m["heads"])
    a2 = Attention(m["C"], m["cad"], m["heads"], m["C"] // m["heads"])
    a1.load_state_dict(case["attn1"]); a2.load_state_dict(case["attn2"])
    p1 = Consistent_AttProcessor(hidden_size=m["C"], cross_attention_dim=None, rank=m["rank"])
    p2 = Consistent_IPAttProcessor(hidden_size=m["C"], cross_attention_dim=m["cad"], rank=m["rank"], scale=m["scale"], num_tokens=4)
    p1.load_state_dict(case["proc1"], strict=True)       # the checkpoint's parameter names load unchanged
    p2.load_state_dict(case["proc2"], strict=True)
    for mod in (a1, a2, p1, p2):
        mod.to("cuda", dtype)
    # set_ip_adapter-style installation through the diffusers protocol (pipline_StableDiffusion_ConsistentID.py:152-174)
    a1.set_processor(p1); a2.set_processor(p2)
    x, ehs = case["x"].to("cuda", dtype), case["ehs"].to("cuda", dtype)
    y1 = a1(x)
    y2 = a2(x, encoder_hidden_states=ehs)
    y2b = a2(x, encoder_hidden_states=ehs)                # cached K/V path
    torch.cuda.synchronize()
    tol = 2e-2 if dtype == torch.float16 else 6e-2       # 16-bit weights + activations vs the fp32 reference outputs
    for got, want in ((y1, case["y_self"]), (y2, case["y_cross"]), (y2b, case["y_cross"])):
        err = (got.float().cpu() - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), (err, want.abs().max().item())
    # the live `scale` attribute (set_scale, :211-214) changes the id branch
    p2.scale = 0.0
    y0 = a2(x, encoder_hidden_states=ehs)
    assert (y0.float() - y2.float()).abs().max().item() > 1e-3


@pytest.mark.gpu
def test_processor_rejects_cpu_tensors():
    from consistentid_b200.processors import Consistent_AttProcessor
    a = Attention(64, None, 2, 32)
    p = Consistent_AttProcessor(hidden_size=64, rank=4)
    with pytest.raises(RuntimeError):
        p(a, torch.randn(1, 16, 64))
