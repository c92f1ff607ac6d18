"""AutoencoderKL (SD-1.5 VAE) encode / decode on the sm_100a kernels.

Mirrors what the diffusers pipelines run around the sampling loop (``vae.encode(...).latent_dist.sample()`` and
``vae.decode``; reference call sites ``src/inference.py:486,566,664,758``; architecture
``outputs/models/denoising/best/vae/config.json``).

B200 specifics: channels-last bf16 activations (the 512x512x128 levels are HBM-bound, so the stream stays in
bf16), convs as tcgen05 implicit GEMMs with fused bias/residual, the nearest-2x upsample + 3x3 conv pair
re-expressed as four 2x2 parity convolutions (2.25x fewer MACs, no upsampled tensor), the VAE-encoder
asymmetric-pad stride-2 conv as parity views, and the single-head 512-wide mid-block attention as
GEMM -> row softmax -> GEMM on the same GEMM kernel (V^T comes out of a role-swapped projection GEMM).
"""
from __future__ import annotations

import torch

from . import ops
from .weights import pack_conv, upsample_parity_weights, vae_param_shapes

bf16, f32 = ops.OPERAND_DTYPE, torch.float32      # bf16 = the build's 16-bit operand dtype (ops.py)


class _VResnet:
    def __init__(self, sd, p, dev):
        g = lambda k: sd[p + k]
        f = lambda k: g(k).to(dev, f32).contiguous()
        self.g1, self.b1, self.g2, self.b2 = f("norm1.weight"), f("norm1.bias"), f("norm2.weight"), f("norm2.bias")
        self.w1, self.cb1 = pack_conv(g("conv1.weight")).to(dev, bf16), f("conv1.bias")
        w2, cb2 = pack_conv(g("conv2.weight")), g("conv2.bias").to(f32)
        self.shortcut = (p + "conv_shortcut.weight") in sd
        if self.shortcut:
            w2 = torch.cat([w2, g("conv_shortcut.weight").flatten(1)], dim=1)
            cb2 = cb2 + g("conv_shortcut.bias").to(f32)
        self.w2, self.cb2 = w2.contiguous().to(dev, bf16), cb2.to(dev).contiguous()


class _VAttn:
    def __init__(self, sd, p, dev):
        g = lambda k: sd[p + k]
        f = lambda k: g(k).to(dev, f32).contiguous()
        self.gn_g, self.gn_b = f("group_norm.weight"), f("group_norm.bias")
        self.w_qk = torch.cat([g("to_q.weight"), g("to_k.weight")], dim=0).to(dev, bf16).contiguous()
        self.b_qk = torch.cat([g("to_q.bias"), g("to_k.bias")], dim=0).to(dev, f32).contiguous()
        self.w_v, self.b_v = g("to_v.weight").to(dev, bf16).contiguous(), f("to_v.bias")
        self.w_o, self.b_o = g("to_out.0.weight").to(dev, bf16).contiguous(), f("to_out.0.bias")
        self.C = self.w_v.shape[0]


class VAEB200:
    scaling_factor = 0.18215

    def __init__(self, state_dict, device: str = "cuda"):
        shapes = vae_param_shapes()
        for k, s in shapes.items():
            if k not in state_dict:
                raise KeyError(f"VAE state dict is missing {k}")
            if tuple(state_dict[k].shape) != tuple(s):
                raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != {s}")
        dev = device
        sd = {k: state_dict[k].to(dev, non_blocking=True) for k in shapes}      # repack on the device (see unet.py)
        self.device = dev
        f = lambda k: sd[k].to(dev, f32).contiguous()
        boc = (128, 256, 512, 512)
        # ---- encoder
        w = torch.zeros((boc[0], 64), dtype=f32, device=dev)
        w[:, :27] = pack_conv(sd["encoder.conv_in.weight"].to(f32))
        self.e_in = (w.to(dev, bf16).contiguous(), f("encoder.conv_in.bias"))
        self.e_down = []
        for i in range(4):
            res = [_VResnet(sd, f"encoder.down_blocks.{i}.resnets.{j}.", dev) for j in range(2)]
            down = None
            if i < 3:
                down = (pack_conv(sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"]).to(dev, bf16),
                        f(f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"))
            self.e_down.append((res, down))
        self.e_mid = (_VResnet(sd, "encoder.mid_block.resnets.0.", dev), _VAttn(sd, "encoder.mid_block.attentions.0.", dev),
                      _VResnet(sd, "encoder.mid_block.resnets.1.", dev))
        self.e_norm = (f("encoder.conv_norm_out.weight"), f("encoder.conv_norm_out.bias"))
        self.e_out = (pack_conv(sd["encoder.conv_out.weight"]).to(dev, bf16), f("encoder.conv_out.bias"))
        self.quant = (sd["quant_conv.weight"].flatten(1).to(dev, f32).contiguous(), f("quant_conv.bias"))
        # ---- decoder
        self.post_quant = (sd["post_quant_conv.weight"].flatten(1).to(dev, f32).contiguous(), f("post_quant_conv.bias"))
        w = torch.zeros((512, 64), dtype=f32, device=dev)
        w[:, :36] = pack_conv(sd["decoder.conv_in.weight"].to(f32))
        self.d_in = (w.to(dev, bf16).contiguous(), f("decoder.conv_in.bias"))
        self.d_mid = (_VResnet(sd, "decoder.mid_block.resnets.0.", dev), _VAttn(sd, "decoder.mid_block.attentions.0.", dev),
                      _VResnet(sd, "decoder.mid_block.resnets.1.", dev))
        self.d_up = []
        for i in range(4):
            res = [_VResnet(sd, f"decoder.up_blocks.{i}.resnets.{j}.", dev) for j in range(3)]
            up = None
            if i < 3:
                wu = sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"]
                up = ([(py, px, wp.to(dev, bf16)) for py, px, wp in upsample_parity_weights(wu)],
                      f(f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"))
            self.d_up.append((res, up))
        self.d_norm = (f("decoder.conv_norm_out.weight"), f("decoder.conv_norm_out.bias"))
        self.d_out = (pack_conv(sd["decoder.conv_out.weight"]).to(dev, bf16), f("decoder.conv_out.bias"))

    # ------------------------------------------------------------------------------------------ blocks
    @staticmethod
    def _resnet(r: _VResnet, x):
        y1, raw = ops.groupnorm(x, r.g1, r.b1, eps=1e-6, silu=True, want_raw=r.shortcut and x.dtype != bf16)
        h, _ = ops.conv2d(y1, r.w1, kh=3, kw=3, pad_t=1, pad_l=1, bias=r.cb1, out_bf16=True)
        y2, _ = ops.groupnorm(h, r.g2, r.b2, eps=1e-6, silu=True)
        if r.shortcut:
            out, _ = ops.conv2d(y2, r.w2, kh=3, kw=3, pad_t=1, pad_l=1, x2=x if raw is None else raw, bias=r.cb2,
                                out_bf16=True)
        else:
            out, _ = ops.conv2d(y2, r.w2, kh=3, kw=3, pad_t=1, pad_l=1, bias=r.cb2, res=x, out_bf16=True)
        return out

    @staticmethod
    def _attn(a: _VAttn, x):
        N, H, W, Cc = x.shape
        T = H * W
        Tp = (T + 7) // 8 * 8                                   # row pitch of the score matrix (16-B aligned rows)
        y, _ = ops.groupnorm(x, a.gn_g, a.gn_b, eps=1e-6, silu=False)
        y2 = y.view(N * T, Cc)
        qk, _ = ops.linear(y2, a.w_qk, images=N, bias=a.b_qk, out_bf16=True)              # [N*T, 2C]
        out = torch.empty_like(x)
        S = torch.empty((T, Tp), dtype=bf16, device=x.device)
        vt = torch.empty((Cc, Tp), dtype=bf16, device=x.device)
        xf = x.view(N * T, Cc)
        of = out.view(N * T, Cc)
        for n in range(N):
            q = qk[n * T:(n + 1) * T, :Cc]
            k = qk[n * T:(n + 1) * T, Cc:]
            yn = y2[n * T:(n + 1) * T]
            # scores = scale * Q K^T  (K plays the weight role: [T, C] K-major)
            ops.conv2d(q.as_strided((1, 1, T, Cc), (0, 0, q.stride(0), 1)), k, w_ld=k.stride(0), scale=Cc ** -0.5,
                       out_bf16=S, out_strides=(0, 0, Tp))
            ops.softmax_rows_(S[:, :T])
            # V^T = Wv Y^T (role swap: the weight matrix is the M operand), bias folded after P V (rows of P sum to 1)
            ops.conv2d(a.w_v.view(1, 1, Cc, Cc), yn, w_ld=yn.stride(0), out_bf16=vt, out_strides=(0, 0, Tp))
            o, _ = ops.conv2d(S.as_strided((1, 1, T, T), (0, 0, Tp, 1)), vt[:, :T], w_ld=Tp, bias=a.b_v, out_bf16=True)
            ops.linear(o.view(T, Cc), a.w_o, bias=a.b_o, res=xf[n * T:(n + 1) * T],
                       out_bf16=of[n * T:(n + 1) * T].view(1, 1, T, Cc))
        return out

    @staticmethod
    def _upsample_conv(up, x):
        N, H, W, Cc = x.shape
        wts, bias = up
        out = torch.empty((N, 2 * H, 2 * W, Cc), dtype=bf16, device=x.device)
        sn, sh, sw = out.stride(0), out.stride(1), out.stride(2)
        for py, px, wp in wts:
            ops.conv2d(x, wp, kh=2, kw=2, pad_t=1 - py, pad_l=1 - px, OH=H, OW=W, bias=bias,
                       out_bf16=out[:, py:, px:], out_strides=(sn, 2 * sh, 2 * sw))
        return out

    # ------------------------------------------------------------------------------------------ API
    def encode_moments(self, image: torch.Tensor) -> torch.Tensor:
        """image: f32 channels-last [N, H, W, 3] in [-1, 1] -> posterior moments f32 [N, H/8, W/8, 8]
        (mean 0..3, logvar 4..7), i.e. quant_conv(encoder(x))."""
        N, H, W, _ = image.shape
        cols = ops.im2col_small(image, N, 3, 1, 1, H, W, 64)
        x, _ = ops.conv2d(cols, self.e_in[0], bias=self.e_in[1], out_bf16=True)
        for res, down in self.e_down:
            for r in res:
                x = self._resnet(r, x)
            if down is not None:
                h, w = x.shape[1], x.shape[2]
                # F.pad(x, (0,1,0,1)) + conv stride 2 pad 0: out = floor((h + 1 - 3) / 2) + 1
                x, _ = ops.conv2d(x, down[0], kh=3, kw=3, stride=2, pad_t=0, pad_l=0, OH=(h - 2) // 2 + 1,
                                  OW=(w - 2) // 2 + 1, bias=down[1], out_bf16=True)
        x = self._resnet(self.e_mid[0], x)
        x = self._attn(self.e_mid[1], x)
        x = self._resnet(self.e_mid[2], x)
        y, _ = ops.groupnorm(x, *self.e_norm, eps=1e-6, silu=True)
        _, m = ops.conv2d(y, self.e_out[0], kh=3, kw=3, pad_t=1, pad_l=1, bias=self.e_out[1], out_f32=True)
        return ops.pointwise_small(m, self.quant[0], self.quant[1])

    def decode(self, latents: torch.Tensor) -> torch.Tensor:
        """latents: f32 channels-last [N, h, w, 4] (still scaled by 0.18215) -> image f32 [N, 8h, 8w, 3]."""
        N, h, w, _ = latents.shape
        z = ops.pointwise_small(latents, self.post_quant[0], self.post_quant[1], scale_in=1.0 / self.scaling_factor)
        cols = ops.im2col_small(z, N, 3, 1, 1, h, w, 64)
        x, _ = ops.conv2d(cols, self.d_in[0], bias=self.d_in[1], out_bf16=True)
        x = self._resnet(self.d_mid[0], x)
        x = self._attn(self.d_mid[1], x)
        x = self._resnet(self.d_mid[2], x)
        for res, up in self.d_up:
            for r in res:
                x = self._resnet(r, x)
            if up is not None:
                x = self._upsample_conv(up, x)
        y, _ = ops.groupnorm(x, *self.d_norm, eps=1e-6, silu=True)
        _, img = ops.conv2d(y, self.d_out[0], kh=3, kw=3, pad_t=1, pad_l=1, bias=self.d_out[1], out_f32=True)
        return img
