"""CPU: host-side logic of the product (no kernels are launched): C-ABI exports, architecture inventory, weight packing."""
import os
import re

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    import ctypes
    from consistentid_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "cidb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cid_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(so, name), f"libcidb200.so does not export {name}"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    assert lib.version() >= 100


def test_param_inventory_matches_oracle():
    from consistentid_b200.arch import attn_processor_names, param_shapes, sd15_spec, sdxl_spec
    from oracle.processors_ref import install_ref_processors
    from oracle.unet_ref import UNet2DConditionRef, sd15_config, sdxl_config
    for spec, cfg, n_params in ((sd15_spec(), sd15_config(), 859_520_964), (sdxl_spec(), sdxl_config(), 2_567_463_684)):
        with torch.device("meta"):
            u = UNet2DConditionRef(cfg)
            install_ref_processors(u, rank=128)
        us, as_ = param_shapes(spec)
        sd = {k: tuple(v.shape) for k, v in u.state_dict().items() if ".processor." not in k}
        assert sd == us
        assert list(u.attn_processors.keys()) == attn_processor_names(spec)      # checkpoint order: down, up, mid
        asd = {k: tuple(v.shape) for k, v in torch.nn.ModuleList(u.attn_processors.values()).state_dict().items()}
        assert asd == as_
        assert sum(torch.Size(s).numel() for s in us.values()) == n_params      # known SD1.5 / SDXL UNet sizes


def test_fold_lora_is_the_processor_math():
    from consistentid_b200.weights import fold_lora
    g = torch.Generator().manual_seed(0)
    w, down, up, x = (torch.randn(s, generator=g) for s in ((48, 32), (8, 32), (48, 8), (5, 32)))
    ref = x @ w.T + 0.7 * (x @ down.T) @ up.T            # attn.to_q(x) + lora_scale * to_q_lora(x)
    assert torch.allclose(x @ fold_lora(w, down, up, 0.7).T, ref, atol=1e-5)


def test_geglu_interleave_roundtrip():
    from consistentid_b200.weights import interleave_geglu
    inner, K, tile = 320, 16, 160
    w, b = torch.randn(2 * inner, K), torch.randn(2 * inner)
    wi, bi = interleave_geglu(w, b, tile)
    x = torch.randn(3, K)
    proj = x @ wi.T + bi
    half = tile // 2
    t = proj.reshape(3, -1, tile)
    out = (t[..., :half] * F.gelu(t[..., half:])).reshape(3, inner)
    v, gt = (x @ w.T + b).chunk(2, -1)
    assert torch.allclose(out, v * F.gelu(gt), atol=1e-5)


def test_conv_weight_packing_is_im2col_order():
    from consistentid_b200.weights import pack_conv3x3
    w = torch.randn(6, 5, 3, 3)
    x = torch.randn(2, 5, 7, 9)
    wp = pack_conv3x3(w, cin_pad=8)
    xp = F.pad(x, (1, 1, 1, 1))
    cols = []
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + 7, kx:kx + 9].permute(0, 2, 3, 1)      # NHWC tap
            cols.append(F.pad(patch, (0, 3)))
    a = torch.cat(cols, -1).reshape(-1, 72)
    ref = F.conv2d(x, w, padding=1).permute(0, 2, 3, 1).reshape(-1, 6)
    assert torch.allclose(a @ wp.T, ref, atol=1e-4)


def test_walk_channel_bookkeeping():
    from consistentid_b200.arch import sd15_spec, sdxl_spec, walk
    ups = {s.name: [[r.cin for r, _ in layers] for kind, _, layers, _ in walk(s) if kind == "up"] for s in (sd15_spec(), sdxl_spec())}
    assert ups["sd15"] == [[2560, 2560, 2560], [2560, 2560, 1920], [1920, 1280, 960], [960, 640, 640]]
    assert ups["sdxl"] == [[2560, 2560, 1920], [1920, 1280, 960], [960, 640, 640]]


def test_product_never_touches_the_oracle_or_a_compiler_stack():
    """The oracle is test infrastructure; the product has no CPU fallback, no Triton / torch.compile path."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = re.compile(r"^\s*(from|import)\s+(oracle|triton|tilelang)\b|torch\.compile\(", re.M)
    for path in glob.glob(os.path.join(root, "consistentid_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not bad.search(src), path
    entry = open(os.path.join(root, "__graft_entry__.py")).read()
    assert "def build" in entry and "def smoke" in entry and "compute_100a" in entry


def test_gn_small_applicability_rule():
    """ops.gn_small_ok mirrors cid_gn_small's contract (include/cidb200.h): 8-channel vectors inside groups, groups inside one source of the
    virtual concat, at most 32768 elements per (sample, group)."""
    from consistentid_b200 import ops
    assert ops.gn_small_ok(1280, 0, 64, 32) and ops.gn_small_ok(1280, 1280, 64, 32)          # the SD1.5 8x8 level
    assert ops.gn_small_ok(1280, 1280, 256, 32)                                              # 16x16: 20480 elements
    assert not ops.gn_small_ok(1280, 640, 256, 32)                                           # 60 channels per group: vectors straddle groups
    assert not ops.gn_small_ok(320, 0, 4096, 32)                                             # slab too large
    assert not ops.gn_small_ok(640, 320, 64, 32) or (960 // 32) % 8 == 0                      # 30 channels per group
    assert not ops.gn_small_ok(328, 0, 64, 32)                                               # C not divisible by the groups
