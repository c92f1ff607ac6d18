"""Weight layout for the sm_100a kernels, parameter-shape tables and seeded random initialisation.

State dicts use diffusers' key names (``down_blocks.0.resnets.0.conv1.weight`` ...) so checkpoints written by
``pipeline.save_pretrained`` (reference ``scripts/train_denoising.py:785``; layout
``outputs/models/<task>/best/{unet,vae}/diffusion_pytorch_model.safetensors``) load unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

GEGLU_TILE = 32           # unit of the GEMM kernel's GEGLU epilogue: 16 value columns | their 16 gate columns


# ------------------------------------------------------------------------------------------------ packing
def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin]: K-major GEMM B operand, taps outer / channels inner."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def interleave_geglu(w: torch.Tensor, b: torch.Tensor | None):
    """Reorder the rows of GEGLU's projection ([2*D, K]: D value rows then D gate rows) so that every
    32-row unit holds 16 value rows followed by their 16 gate rows."""
    D = w.shape[0] // 2
    half = GEGLU_TILE // 2
    assert D % half == 0
    idx = torch.arange(D, device=w.device).view(D // half, half)
    perm = torch.cat([idx, idx + D], dim=1).reshape(-1)
    return w[perm].contiguous(), (None if b is None else b[perm].contiguous())


def upsample_parity_weights(w: torch.Tensor) -> list:
    """Nearest-2x upsample followed by a 3x3/pad-1 conv == four 2x2 convs on the low-res input, one per output
    parity (py, px).  Row taps collapse as: py=0 -> {kh0 | kh1+kh2} at rows (j-1, j); py=1 -> {kh0+kh1 | kh2}
    at rows (j, j+1); same for columns.  Returns [(py, px, packed [Cout, 4*Cin])], summed in fp32."""
    w = w.float()
    rows = {0: [w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1: [w[:, :, 0] + w[:, :, 1], w[:, :, 2]]}
    out = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for r in rows[py]:                       # r: [Cout, Cin, 3(kw)]
                cols = {0: [r[:, :, 0], r[:, :, 1] + r[:, :, 2]], 1: [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]}[px]
                taps.extend(cols)                    # order (row tap, col tap)
            out.append((py, px, torch.cat(taps, dim=1).contiguous()))
    return out


# ------------------------------------------------------------------------------------------------ shape tables
def _resnet(sd, p, cin, cout, temb):
    sd[p + "norm1.weight"] = (cin,); sd[p + "norm1.bias"] = (cin,)
    sd[p + "conv1.weight"] = (cout, cin, 3, 3); sd[p + "conv1.bias"] = (cout,)
    if temb:
        sd[p + "time_emb_proj.weight"] = (cout, temb); sd[p + "time_emb_proj.bias"] = (cout,)
    sd[p + "norm2.weight"] = (cout,); sd[p + "norm2.bias"] = (cout,)
    sd[p + "conv2.weight"] = (cout, cout, 3, 3); sd[p + "conv2.bias"] = (cout,)
    if cin != cout:
        sd[p + "conv_shortcut.weight"] = (cout, cin, 1, 1); sd[p + "conv_shortcut.bias"] = (cout,)


def _transformer(sd, p, c, ctx):
    sd[p + "norm.weight"] = (c,); sd[p + "norm.bias"] = (c,)
    sd[p + "proj_in.weight"] = (c, c, 1, 1); sd[p + "proj_in.bias"] = (c,)
    t = p + "transformer_blocks.0."
    for n in ("norm1", "norm2", "norm3"):
        sd[t + n + ".weight"] = (c,); sd[t + n + ".bias"] = (c,)
    for a, kv in (("attn1", c), ("attn2", ctx)):
        sd[t + a + ".to_q.weight"] = (c, c)
        sd[t + a + ".to_k.weight"] = (c, kv)
        sd[t + a + ".to_v.weight"] = (c, kv)
        sd[t + a + ".to_out.0.weight"] = (c, c); sd[t + a + ".to_out.0.bias"] = (c,)
    sd[t + "ff.net.0.proj.weight"] = (8 * c, c); sd[t + "ff.net.0.proj.bias"] = (8 * c,)
    sd[t + "ff.net.2.weight"] = (c, 4 * c); sd[t + "ff.net.2.bias"] = (c,)
    sd[p + "proj_out.weight"] = (c, c, 1, 1); sd[p + "proj_out.bias"] = (c,)


def unet_param_shapes(in_channels: int = 4, out_channels: int = 4, boc=(320, 640, 1280, 1280), layers: int = 2,
                      ctx: int = 768, attn=(True, True, True, False)) -> "OrderedDict[str, tuple]":
    """Every parameter of UNet2DConditionModel (SD-1.5 family) by diffusers key.
    Follows ``outputs/models/denoising/best/unet/config.json``; total 859,520,964 for in_channels=4
    (the count the reference logged, ``outputs/models/colorization/training_colorization.log:30``)."""
    sd: OrderedDict = OrderedDict()
    temb = boc[0] * 4
    sd["conv_in.weight"] = (boc[0], in_channels, 3, 3); sd["conv_in.bias"] = (boc[0],)
    sd["time_embedding.linear_1.weight"] = (temb, boc[0]); sd["time_embedding.linear_1.bias"] = (temb,)
    sd["time_embedding.linear_2.weight"] = (temb, temb); sd["time_embedding.linear_2.bias"] = (temb,)
    ch = boc[0]
    for i, co in enumerate(boc):
        for j in range(layers):
            _resnet(sd, f"down_blocks.{i}.resnets.{j}.", ch if j == 0 else co, co, temb)
            if attn[i]:
                _transformer(sd, f"down_blocks.{i}.attentions.{j}.", co, ctx)
        if i != len(boc) - 1:
            sd[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (co, co, 3, 3)
            sd[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (co,)
        ch = co
    c = boc[-1]
    _resnet(sd, "mid_block.resnets.0.", c, c, temb)
    _transformer(sd, "mid_block.attentions.0.", c, ctx)
    _resnet(sd, "mid_block.resnets.1.", c, c, temb)
    rev = list(reversed(boc))
    rattn = list(reversed(attn))
    prev = rev[0]
    for i, co in enumerate(rev):
        cin_blk = rev[min(i + 1, len(boc) - 1)]
        for j in range(layers + 1):
            skip = cin_blk if j == layers else co
            rin = prev if j == 0 else co
            _resnet(sd, f"up_blocks.{i}.resnets.{j}.", rin + skip, co, temb)
            if rattn[i]:
                _transformer(sd, f"up_blocks.{i}.attentions.{j}.", co, ctx)
        if i != len(boc) - 1:
            sd[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            sd[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
        prev = co
    sd["conv_norm_out.weight"] = (boc[0],); sd["conv_norm_out.bias"] = (boc[0],)
    sd["conv_out.weight"] = (out_channels, boc[0], 3, 3); sd["conv_out.bias"] = (out_channels,)
    return sd


def _vae_mid(sd, p, c):
    _resnet(sd, p + "resnets.0.", c, c, None)
    a = p + "attentions.0."
    sd[a + "group_norm.weight"] = (c,); sd[a + "group_norm.bias"] = (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        sd[a + n + ".weight"] = (c, c); sd[a + n + ".bias"] = (c,)
    _resnet(sd, p + "resnets.1.", c, c, None)


def vae_param_shapes(in_channels=3, out_channels=3, latent=4, boc=(128, 256, 512, 512), layers=2):
    """Every parameter of AutoencoderKL by diffusers key (``outputs/models/denoising/best/vae/config.json``);
    total 83,653,863."""
    sd: OrderedDict = OrderedDict()
    sd["encoder.conv_in.weight"] = (boc[0], in_channels, 3, 3); sd["encoder.conv_in.bias"] = (boc[0],)
    ch = boc[0]
    for i, co in enumerate(boc):
        for j in range(layers):
            _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}.", ch if j == 0 else co, co, None)
        if i != len(boc) - 1:
            sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (co, co, 3, 3)
            sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (co,)
        ch = co
    _vae_mid(sd, "encoder.mid_block.", boc[-1])
    sd["encoder.conv_norm_out.weight"] = (boc[-1],); sd["encoder.conv_norm_out.bias"] = (boc[-1],)
    sd["encoder.conv_out.weight"] = (2 * latent, boc[-1], 3, 3); sd["encoder.conv_out.bias"] = (2 * latent,)
    rev = list(reversed(boc))
    sd["decoder.conv_in.weight"] = (rev[0], latent, 3, 3); sd["decoder.conv_in.bias"] = (rev[0],)
    _vae_mid(sd, "decoder.mid_block.", rev[0])
    ch = rev[0]
    for i, co in enumerate(rev):
        for j in range(layers + 1):
            _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}.", ch if j == 0 else co, co, None)
        if i != len(rev) - 1:
            sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
        ch = co
    sd["decoder.conv_norm_out.weight"] = (boc[0],); sd["decoder.conv_norm_out.bias"] = (boc[0],)
    sd["decoder.conv_out.weight"] = (out_channels, boc[0], 3, 3); sd["decoder.conv_out.bias"] = (out_channels,)
    sd["quant_conv.weight"] = (2 * latent, 2 * latent, 1, 1); sd["quant_conv.bias"] = (2 * latent,)
    sd["post_quant_conv.weight"] = (latent, latent, 1, 1); sd["post_quant_conv.bias"] = (latent,)
    return sd


# ------------------------------------------------------------------------------------------------ random init
def random_state_dict(shapes: "OrderedDict[str, tuple]", seed: int, *, gain: float = 1.0,
                      out_gain: float = 1.0, device: str = "cpu") -> "OrderedDict[str, torch.Tensor]":
    """Seeded, variance-preserving random initialisation shared verbatim by the oracle and the CUDA path.

    * matrices / conv kernels: N(0, gain^2 / fan_in); norm scales 1 + 0.1 N(0,1); all biases 0.05 N(0,1);
    * every value is rounded to bf16 and stored as fp32, so the fp32 oracle and the bf16 tensor-core path use
      numerically identical weights (the parity gates then measure arithmetic, not weight quantisation);
    * ``out_gain`` scales the network's last convolution (``conv_out``) so outputs stay O(1).
    There are no pretrained weights offline; real checkpoints load through the same keys.
    ``device``: a CUDA device draws the numbers there (Philox; seconds faster than 860 M values on the host -- the sweep's
    cold start), the default draws on the CPU generator (the values the parity tests share with the oracle).
    """
    on_gpu = torch.device(device).type == "cuda"
    g = torch.Generator(device=device if on_gpu else "cpu").manual_seed(seed)
    gen_dev = device if on_gpu else "cpu"
    sd: OrderedDict = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith(".weight") and len(shape) >= 2:
            fan_in = math.prod(shape[1:])
            s = gain / math.sqrt(fan_in)
            if name in ("conv_out.weight", "decoder.conv_out.weight", "encoder.conv_out.weight"):
                s *= out_gain
            t = torch.randn(shape, generator=g, device=gen_dev) * s
        elif name.endswith(".weight"):           # norm scale
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=gen_dev)
        else:
            t = 0.05 * torch.randn(shape, generator=g, device=gen_dev)
        sd[name] = t.to(torch.bfloat16).to(torch.float32).to(device)
    return sd
