"""fp32 PyTorch restatement of diffusers' ``AutoencoderKL`` (SD-1.5 VAE).

Test infrastructure only (see ``oracle/__init__.py``).

Architecture pinned by
``/root/reference/outputs/models/denoising/best/vae/config.json:1-38``; used by the
reference through ``vae.encode(...).latent_dist.sample()`` / ``vae.decode`` inside
the diffusers pipelines called at ``src/inference.py:486,566,664,758`` (and directly
in ``scripts/train_denoising.py:629-630,691``).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import ResnetBlock2D, Downsample2D, Upsample2D


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @classmethod
    def from_json(cls, d: dict) -> "VAEConfig":
        keys = cls.__dataclass_fields__.keys()
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items() if k in keys}
        return cls(**kw)


class VAEAttention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block (deprecated-attn-block form)."""

    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6, affine=True)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        res = x
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(B, C, H, W) + res


class VAEMidBlock(nn.Module):
    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, groups, 1e-6),
                                      ResnetBlock2D(ch, ch, None, groups, 1e-6)])
        self.attentions = nn.ModuleList([VAEAttention(ch, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, n_layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, None, groups, 1e-6) for i in range(n_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch, padding=0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, n_layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, None, groups, 1e-6) for i in range(n_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownEncoderBlock2D(in_ch, out_ch, cfg.layers_per_block, g,
                                                       add_down=i != len(boc) - 1))
        self.mid_block = VAEMidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = VAEMidBlock(rev[0], g)
        self.up_blocks = nn.ModuleList()
        out_ch = rev[0]
        for i in range(len(rev)):
            in_ch, out_ch = out_ch, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(in_ch, out_ch, cfg.layers_per_block + 1, g,
                                                   add_up=i != len(rev) - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise: torch.Tensor | None = None) -> torch.Tensor:
        if noise is None:
            noise = randn_tensor(self.mean.shape, generator, self.mean.device, self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


def randn_tensor(shape, generator, device, dtype):
    """diffusers ``randn_tensor``: a CPU generator samples on CPU and moves; else on-device."""
    device = torch.device(device)
    gen_dev = generator.device.type if generator is not None else device.type
    if gen_dev != device.type and gen_dev == "cpu":
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def encode(self, x) -> DiagonalGaussianDistribution:
        return DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))

    def decode(self, z) -> torch.Tensor:
        return self.decoder(self.post_quant_conv(z))
