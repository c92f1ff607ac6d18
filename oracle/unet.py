"""fp32 PyTorch restatement of diffusers' ``UNet2DConditionModel`` (SD-1.5 family).

Test infrastructure only (see ``oracle/__init__.py``).

Follows the architecture pinned by
``/root/reference/outputs/models/denoising/best/unet/config.json:1-68`` (4-channel)
and ``outputs/models/inpainting/best/unet/config.json:37`` (``in_channels: 9``),
which is what ``src/inference.py:486,566,664,758`` ends up executing through the
diffusers pipelines.  Module and parameter names match diffusers' state-dict keys.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    # defaults == outputs/models/denoising/best/unet/config.json
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    down_block_types: tuple = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                               "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: tuple = ("UpBlock2D", "CrossAttnUpBlock2D",
                             "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    attention_head_dim: int = 8        # SD-1.5: this field is the HEAD COUNT
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    sample_size: int = 64

    @classmethod
    def from_json(cls, d: dict) -> "UNetConfig":
        keys = cls.__dataclass_fields__.keys()
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items() if k in keys}
        return cls(**kw)


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool,
                       freq_shift: float, max_period: int = 10000) -> torch.Tensor:
    """diffusers ``get_timestep_embedding`` (embeddings.py)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_ch: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_ch, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, temb_ch: int | None, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_ch, eps=eps)
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, out_ch) if temb_ch is not None else None
        self.norm2 = nn.GroupNorm(groups, out_ch, eps=eps)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers ``Attention`` with ``AttnProcessor2_0`` (SDPA, no mask, dropout 0)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int,
                 cross_attention_dim: int | None = None, bias: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        h = self.heads
        q = q.view(B, N, h, -1).transpose(1, 2)
        k = k.view(B, ctx.shape[1], h, -1).transpose(1, 2)
        v = v.view(B, ctx.shape[1], h, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(B, N, -1)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)          # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    """``use_linear_projection=False`` => 1x1 conv proj_in/proj_out; GroupNorm eps 1e-6."""

    def __init__(self, heads: int, dim_head: int, in_ch: int, cross_dim: int, groups: int):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_ch, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_ch, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_dim)])
        self.proj_out = nn.Conv2d(inner, in_ch, 1)

    def forward(self, x, context):
        B, C, H, W = x.shape
        res = x
        h = self.proj_in(self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, -1)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        h = h.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(h) + res


class Downsample2D(nn.Module):
    def __init__(self, ch: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x, output_size=None):
        # diffusers: explicit `size=` when the UNet input is not divisible by 2**num_upsamplers
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, in_ch, out_ch, temb_ch, n_layers, groups, eps, heads, cross_dim,
                 has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(n_layers):
            self.resnets.append(ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, out_ch // heads, out_ch, cross_dim, groups))
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch, 1)]) if add_down else None

    def forward(self, x, temb, context):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, groups, eps, heads, cross_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, groups, eps),
                                      ResnetBlock2D(ch, ch, temb_ch, groups, eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, ch, cross_dim, groups)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, in_ch, out_ch, prev_ch, temb_ch, n_layers, groups, eps, heads, cross_dim,
                 has_attn, add_up):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(n_layers):
            skip_ch = in_ch if i == n_layers - 1 else out_ch
            res_in = prev_ch if i == 0 else out_ch
            self.resnets.append(ResnetBlock2D(res_in + skip_ch, out_ch, temb_ch, groups, eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, out_ch // heads, out_ch, cross_dim, groups))
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_up else None

    def forward(self, x, skips, temb, context, forward_upsample_size=False):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
        if self.upsamplers is not None:
            size = tuple(skips[-1].shape[2:]) if forward_upsample_size else None
            x = self.upsamplers[0](x, size)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        temb_ch = boc[0] * 4
        heads, cd, g, eps = cfg.attention_head_dim, cfg.cross_attention_dim, cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, cfg.layers_per_block, g, eps, heads, cd,
                                              has_attn=t.startswith("CrossAttn"),
                                              add_down=i != len(boc) - 1))
        self.mid_block = MidBlock(boc[-1], temb_ch, g, eps, heads, cd)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, t in enumerate(cfg.up_block_types):
            prev_ch, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev_ch, temb_ch, cfg.layers_per_block + 1, g, eps,
                                          heads, cd, has_attn=t.startswith("CrossAttn"),
                                          add_up=i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        t = torch.as_tensor(timestep, device=sample.device)
        if t.ndim == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        t_emb = timestep_embedding(t, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        temb = self.time_embedding(t_emb.to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states)
        # default_overall_up_factor = 2 ** 3: inputs not divisible by 8 make every upsampler use the skip's size
        fus = any(d % 8 != 0 for d in sample.shape[-2:])
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states, fus)
        x = F.silu(self.conv_norm_out(x))
        return self.conv_out(x)
