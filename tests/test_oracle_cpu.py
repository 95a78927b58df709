"""CPU: pin the oracle's restatement of the reference processors against (a) golden vectors generated from the verbatim
reference (tests/golden/make_golden.py) and (b) the verbatim reference itself when /root/reference is present."""
import os
import sys

import pytest
import torch

from oracle.processors_ref import ConsistentAttnRef, ConsistentIPAttnRef
from oracle.unet_ref import Attention

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "processors_golden.pt")


def _build(case):
    m = case["meta"]
    a1, a2 = Attention(m["C"], None, m["heads"], m["C"] // m["heads"]), Attention(m["C"], m["cad"], m["heads"], m["C"] // m["heads"])
    a1.load_state_dict(case["attn1"]); a2.load_state_dict(case["attn2"])
    p1 = ConsistentAttnRef(hidden_size=m["C"], cross_attention_dim=None, rank=m["rank"])
    p2 = ConsistentIPAttnRef(hidden_size=m["C"], cross_attention_dim=m["cad"], rank=m["rank"], scale=m["scale"], num_tokens=4)
    p1.load_state_dict(case["proc1"], strict=True)      # same parameter names as the reference processors
    p2.load_state_dict(case["proc2"], strict=True)
    return a1, a2, p1, p2


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_matches_reference_golden(idx):
    case = torch.load(GOLDEN)[idx]
    a1, a2, p1, p2 = _build(case)
    with torch.no_grad():
        y1 = p1(a1, case["x"])
        y2 = p2(a2, case["x"], encoder_hidden_states=case["ehs"])
    assert torch.allclose(y1, case["y_self"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(y2, case["y_cross"], rtol=1e-5, atol=1e-5)


def test_oracle_4d_input_path():
    case = torch.load(GOLDEN)[0]
    a1, a2, p1, p2 = _build(case)
    B, N, C = case["x"].shape
    x4 = case["x"].transpose(1, 2).reshape(B, C, 8, N // 8)
    with torch.no_grad():
        y = p2(a2, x4, encoder_hidden_states=case["ehs"])
    assert torch.allclose(y.reshape(B, C, N).transpose(1, 2), case["y_cross"], rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not os.path.exists("/root/reference/attention.py"), reason="reference tree only exists in the build container")
def test_oracle_matches_verbatim_reference_full_unet():
    root = os.path.dirname(os.path.dirname(__file__))
    sys.path.insert(0, os.path.join(root, "oracle", "diffusers_shim"))
    sys.path.insert(0, "/root/reference")
    import attention as ref_attention
    from oracle import synth
    from oracle.unet_ref import tiny_config
    cfg = tiny_config("sd15")
    cfg.sample_size = 16
    unet = synth.build_ref_unet(cfg, rank=8)
    mine = unet.attn_processors
    theirs = {}
    for name, p in mine.items():
        if isinstance(p, ConsistentIPAttnRef):
            rp = ref_attention.Consistent_IPAttProcessor(p.hidden_size, p.cross_attention_dim, rank=8, num_tokens=4)
        else:
            rp = ref_attention.Consistent_AttProcessor(p.to_q_lora.down.in_features, None, rank=8)
        rp.load_state_dict(p.state_dict(), strict=True)
        theirs[name] = rp
    null, aug, _ = synth.synth_prompts(cfg.cross_attention_dim)
    x = synth.synth_latents(2, 16, 16)
    ehs = torch.cat([null, aug])
    with torch.no_grad():
        y_mine = unet(x, torch.tensor(500), ehs).sample
        unet.set_attn_processor(theirs)
        y_ref = unet(x, torch.tensor(500), ehs).sample
    assert torch.allclose(y_mine, y_ref, rtol=1e-4, atol=1e-5)


# ---- independent cross-check of the (otherwise unpinned) diffusers restatement: TVM's relax port of diffusers' get_timestep_embedding
TVM_OP = "/opt/prime-rl/.venv/lib/python3.12/site-packages/tilelang/3rdparty/tvm/python/tvm/relax/frontend/nn/op.py"


def _run_tvm_port_with_numpy(timesteps, dim, **kw):
    """Execute the SOURCE of tvm.relax.frontend.nn.op.get_timestep_embedding (a third-party port of the diffusers function)
    with its relax ops bound to numpy: the port's own arithmetic graph, evaluated without a TVM runtime."""
    import ast, math, types
    import numpy as np
    src = open(TVM_OP).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_timestep_embedding")
    node.returns = None
    for a in node.args.args:
        a.annotation = None
    code = compile(ast.Module(body=[node], type_ignores=[]), TVM_OP, "exec")
    f32 = lambda x: np.asarray(x, dtype=np.float32)
    op = types.SimpleNamespace(
        astype=lambda x, dt: f32(x), arange=lambda start, end, dtype: np.arange(start, end, dtype=np.float32), exp=lambda x: np.exp(f32(x)),
        expand_dims=lambda x, ax: np.expand_dims(x, ax), concat=lambda xs, axis: np.concatenate(xs, axis=axis), cos=lambda x: np.cos(f32(x)),
        sin=lambda x: np.sin(f32(x)), nn=types.SimpleNamespace(pad=lambda x, p: np.pad(x, ((p[2], p[3]), (p[0], p[1])))))
    ns = dict(math=math, _op=op, rx=types.SimpleNamespace(const=lambda v, dt: np.float32(v)), get_default_dtype=lambda: "float32",
              wrap_nested=lambda e, name: e, Tensor=object)
    exec(code, ns)
    return ns["get_timestep_embedding"](types.SimpleNamespace(_expr=np.asarray(timesteps)), dim, **kw)


@pytest.mark.skipif(not os.path.exists(TVM_OP), reason="TVM sources not in this image")
@pytest.mark.parametrize("dim,flip,shift", [(320, True, 0.0), (256, True, 0.0), (320, False, 1.0)])
def test_timestep_embedding_matches_the_tvm_port_of_diffusers(dim, flip, shift):
    """unet_ref.get_timestep_embedding (diffusers 0.23 restated, parity otherwise unpinned) against an independent port of the same function."""
    from oracle.unet_ref import get_timestep_embedding
    t = torch.tensor([0, 1, 33, 500, 961, 999])
    ours = get_timestep_embedding(t, dim, flip_sin_to_cos=flip, downscale_freq_shift=shift).numpy()
    theirs = _run_tvm_port_with_numpy(t.numpy(), dim, flip_sin_to_cos=flip, downscale_freq_shift=shift)
    assert ours.shape == theirs.shape
    assert abs(ours - theirs).max() < 2e-4      # fp32 sin/cos of arguments up to ~1e3 rad: ulp-level differences of the product; a layout or formula mismatch would be O(1)
