"""Architecture description of the UNets the ConsistentID pipelines drive (diffusers 0.23 ``UNet2DConditionModel`` config
fields) and the parameter inventory (diffusers state_dict names + the reference's ``adapter_modules``
positional keys, pipline_StableDiffusion_ConsistentID.py:143-144)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple


@dataclass
class UNetSpec:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    num_attention_heads: Tuple[int, ...] = (8, 8, 8, 8)      # diffusers calls this "attention_head_dim"
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: Optional[int] = None
    projection_class_embeddings_input_dim: Optional[int] = None
    sample_size: int = 64
    name: str = "sd15"

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @classmethod
    def from_config(cls, cfg):
        """Accepts any object/dict with diffusers-style fields (e.g. ``unet.config``)."""
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        heads = get("num_attention_heads") or get("attention_head_dim")
        boc = tuple(get("block_out_channels"))
        if isinstance(heads, int):
            heads = (heads,) * len(boc)
        tl = get("transformer_layers_per_block", 1)
        if isinstance(tl, int):
            tl = (tl,) * len(boc)
        return cls(in_channels=get("in_channels", 4), out_channels=get("out_channels", 4), block_out_channels=boc,
                   down_block_types=tuple(get("down_block_types")), up_block_types=tuple(get("up_block_types")),
                   layers_per_block=get("layers_per_block", 2), transformer_layers_per_block=tuple(tl),
                   num_attention_heads=tuple(heads), cross_attention_dim=get("cross_attention_dim"),
                   norm_num_groups=get("norm_num_groups", 32), norm_eps=get("norm_eps", 1e-5),
                   use_linear_projection=bool(get("use_linear_projection", False)),
                   addition_embed_type=get("addition_embed_type"), addition_time_embed_dim=get("addition_time_embed_dim"),
                   projection_class_embeddings_input_dim=get("projection_class_embeddings_input_dim"),
                   sample_size=get("sample_size", 64), name=get("name", "unet"))


def sd15_spec(in_channels=4) -> UNetSpec:
    return UNetSpec(in_channels=in_channels, name="sd15")


def sdxl_spec() -> UNetSpec:
    return UNetSpec(block_out_channels=(320, 640, 1280),
                    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
                    transformer_layers_per_block=(1, 2, 10), num_attention_heads=(5, 10, 20), cross_attention_dim=2048,
                    use_linear_projection=True, addition_embed_type="text_time", addition_time_embed_dim=256,
                    projection_class_embeddings_input_dim=2816, sample_size=128, name="sdxl")


# ---------------------------------------------------------------------------------------------- structure walk
@dataclass
class ResnetDesc:
    name: str
    cin: int
    cout: int
    skip_ch: int = 0        # channels of the skip tensor concatenated in front of this resnet (up path)


@dataclass
class TransformerDesc:
    name: str
    channels: int
    heads: int
    layers: int


def walk(spec: UNetSpec):
    """Yield the block structure in execution order:
    ('down', i, [(ResnetDesc, TransformerDesc|None)...], has_downsample), ('mid', ...), ('up', i, [...], has_upsample)."""
    boc = spec.block_out_channels
    out = []
    ch = boc[0]
    skip_chs = [ch]
    for i, t in enumerate(spec.down_block_types):
        cin, ch = ch, boc[i]
        layers = []
        for j in range(spec.layers_per_block):
            r = ResnetDesc(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else ch, ch)
            tf = None
            if t.startswith("CrossAttn"):
                tf = TransformerDesc(f"down_blocks.{i}.attentions.{j}", ch, spec.num_attention_heads[i], spec.transformer_layers_per_block[i])
            layers.append((r, tf))
            skip_chs.append(ch)
        has_ds = i != len(boc) - 1
        if has_ds:
            skip_chs.append(ch)
        out.append(("down", i, layers, has_ds))
    c = boc[-1]
    out.append(("mid", 0, [(ResnetDesc("mid_block.resnets.0", c, c), TransformerDesc("mid_block.attentions.0", c, spec.num_attention_heads[-1], spec.transformer_layers_per_block[-1])),
                           (ResnetDesc("mid_block.resnets.1", c, c), None)], False))
    rev = list(reversed(boc))
    rev_heads = list(reversed(spec.num_attention_heads))
    rev_tf = list(reversed(spec.transformer_layers_per_block))
    ch = rev[0]
    for i, t in enumerate(spec.up_block_types):
        prev, ch = ch, rev[i]
        layers = []
        n = spec.layers_per_block + 1
        for j in range(n):
            skip = skip_chs.pop()
            rin = prev if j == 0 else ch
            r = ResnetDesc(f"up_blocks.{i}.resnets.{j}", rin + skip, ch, skip_ch=skip)
            tf = None
            if t.startswith("CrossAttn"):
                tf = TransformerDesc(f"up_blocks.{i}.attentions.{j}", ch, rev_heads[i], rev_tf[i])
            layers.append((r, tf))
        out.append(("up", i, layers, i != len(boc) - 1))
    assert not skip_chs
    return out


def attn_processor_names(spec: UNetSpec):
    """Order of ``unet.attn_processors`` in diffusers: module registration order down_blocks -> up_blocks -> mid_block,
    attn1 before attn2; the reference checkpoint's ``adapter_modules`` keys are positions in this list."""
    names = []
    blocks = walk(spec)
    ordered = [b for b in blocks if b[0] == "down"] + [b for b in blocks if b[0] == "up"] + [b for b in blocks if b[0] == "mid"]
    for _, _, layers, _ in ordered:
        for _, tf in layers:
            if tf is None:
                continue
            for k in range(tf.layers):
                names.append(f"{tf.name}.transformer_blocks.{k}.attn1.processor")
                names.append(f"{tf.name}.transformer_blocks.{k}.attn2.processor")
    return names


def param_shapes(spec: UNetSpec, rank=128, num_tokens=4):
    """(unet_shapes, adapter_shapes): dicts name -> shape for the diffusers UNet state_dict and for the positional
    ``adapter_modules`` state_dict."""
    u = {}
    boc = spec.block_out_channels
    T = spec.time_embed_dim
    g = lambda n, *s: u.__setitem__(n, tuple(s))
    g("conv_in.weight", boc[0], spec.in_channels, 3, 3); g("conv_in.bias", boc[0])
    g("time_embedding.linear_1.weight", T, boc[0]); g("time_embedding.linear_1.bias", T)
    g("time_embedding.linear_2.weight", T, T); g("time_embedding.linear_2.bias", T)
    if spec.addition_embed_type == "text_time":
        P = spec.projection_class_embeddings_input_dim
        g("add_embedding.linear_1.weight", T, P); g("add_embedding.linear_1.bias", T)
        g("add_embedding.linear_2.weight", T, T); g("add_embedding.linear_2.bias", T)

    def resnet(r: ResnetDesc):
        n = r.name
        g(f"{n}.norm1.weight", r.cin); g(f"{n}.norm1.bias", r.cin)
        g(f"{n}.conv1.weight", r.cout, r.cin, 3, 3); g(f"{n}.conv1.bias", r.cout)
        g(f"{n}.time_emb_proj.weight", r.cout, T); g(f"{n}.time_emb_proj.bias", r.cout)
        g(f"{n}.norm2.weight", r.cout); g(f"{n}.norm2.bias", r.cout)
        g(f"{n}.conv2.weight", r.cout, r.cout, 3, 3); g(f"{n}.conv2.bias", r.cout)
        if r.cin != r.cout:
            g(f"{n}.conv_shortcut.weight", r.cout, r.cin, 1, 1); g(f"{n}.conv_shortcut.bias", r.cout)

    def transformer(t: TransformerDesc):
        n, C = t.name, t.channels
        g(f"{n}.norm.weight", C); g(f"{n}.norm.bias", C)
        if spec.use_linear_projection:
            g(f"{n}.proj_in.weight", C, C); g(f"{n}.proj_out.weight", C, C)
        else:
            g(f"{n}.proj_in.weight", C, C, 1, 1); g(f"{n}.proj_out.weight", C, C, 1, 1)
        g(f"{n}.proj_in.bias", C); g(f"{n}.proj_out.bias", C)
        for k in range(t.layers):
            b = f"{n}.transformer_blocks.{k}"
            for i in (1, 2, 3):
                g(f"{b}.norm{i}.weight", C); g(f"{b}.norm{i}.bias", C)
            for a, kv in (("attn1", C), ("attn2", spec.cross_attention_dim)):
                g(f"{b}.{a}.to_q.weight", C, C); g(f"{b}.{a}.to_k.weight", C, kv); g(f"{b}.{a}.to_v.weight", C, kv)
                g(f"{b}.{a}.to_out.0.weight", C, C); g(f"{b}.{a}.to_out.0.bias", C)
            g(f"{b}.ff.net.0.proj.weight", 8 * C, C); g(f"{b}.ff.net.0.proj.bias", 8 * C)
            g(f"{b}.ff.net.2.weight", C, 4 * C); g(f"{b}.ff.net.2.bias", C)

    for kind, i, layers, has_sampler in walk(spec):
        for r, tf in layers:
            resnet(r)
            if tf is not None:
                transformer(tf)
        if has_sampler:
            c = layers[-1][0].cout
            nm = f"down_blocks.{i}.downsamplers.0.conv" if kind == "down" else f"up_blocks.{i}.upsamplers.0.conv"
            g(f"{nm}.weight", c, c, 3, 3); g(f"{nm}.bias", c)
    g("conv_norm_out.weight", boc[0]); g("conv_norm_out.bias", boc[0])
    g("conv_out.weight", spec.out_channels, boc[0], 3, 3); g("conv_out.bias", spec.out_channels)

    a = {}
    hidden_of = {}
    for kind, i, layers, _ in walk(spec):
        for _, tf in layers:
            if tf is not None:
                hidden_of[tf.name] = tf.channels
    for pos, name in enumerate(attn_processor_names(spec)):
        tfname = name.split(".transformer_blocks.")[0]
        C = hidden_of[tfname]
        kv = C if ".attn1." in name else spec.cross_attention_dim
        for proj, cin in (("to_q", C), ("to_k", kv), ("to_v", kv), ("to_out", C)):
            a[f"{pos}.{proj}_lora.down.weight"] = (rank, cin)
            a[f"{pos}.{proj}_lora.up.weight"] = (C, rank)
        if ".attn2." in name:
            a[f"{pos}.to_k_ip.weight"] = (C, kv)
            a[f"{pos}.to_v_ip.weight"] = (C, kv)
    return u, a


def synth_state_dicts(spec: UNetSpec, device, dtype, seed=1234, rank=128):
    """Random-init weights of the exact architecture, generated directly on ``device`` (bench / smoke data;
    there are no pretrained weights in this environment)."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    ushapes, ashapes = param_shapes(spec, rank)

    def make(name, shape):
        if name.endswith(".bias") and len(shape) == 1 and ".norm" not in name and "conv_norm_out" not in name:
            return (0.02 * torch.randn(shape, generator=gen, device=device)).to(dtype)
        if ".norm" in name or "conv_norm_out" in name:
            base = 1.0 if name.endswith("weight") else 0.0
            return (base + 0.05 * torch.randn(shape, generator=gen, device=device)).to(dtype)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        std = fan_in ** -0.5
        if "lora.up" in name:
            std = 0.02
        if "lora.down" in name:
            std = 1.0 / rank
        return (std * torch.randn(shape, generator=gen, device=device)).to(dtype)

    return {n: make(n, s) for n, s in ushapes.items()}, {n: make(n, s) for n, s in ashapes.items()}
