"""UNet2DConditionModel (SD-1.5 family) forward pass on the sm_100a kernels.

Mirrors the module tree the reference executes through diffusers at ``src/inference.py:486,566,664,758``
(architecture: ``outputs/models/denoising/best/unet/config.json``), re-scheduled for B200:

* activations are channels-last; the residual stream is fp32, every GEMM operand is bf16;
* every conv / linear is one ``rg_conv2d`` launch (tcgen05 implicit GEMM) with bias, time-embedding add,
  residual add, GEGLU and dtype conversion fused into its epilogue;
* the 1x1 ``conv_shortcut`` of a resnet is folded into ``conv2`` as extra K blocks (one GEMM, no extra pass);
* skip-connection concats are never materialised in fp32: GroupNorm reads both halves;
* q/k/v projections are one GEMM; cross-attention K/V are computed once per prompt (``prepare_context``);
* the 22 ``time_emb_proj`` linears are one GEMM per step.
Every call goes to librestoragen.so; nothing here computes with torch.
"""
from __future__ import annotations

import math

import torch

from . import ops
from ._lib import RG_ACT_GEGLU, RG_ACT_SILU
from .weights import interleave_geglu, pack_conv, unet_param_shapes, upsample_parity_weights

bf16, f32, f16 = ops.OPERAND_DTYPE, torch.float32, torch.float16      # bf16 = the build's 16-bit operand dtype (ops.py)


class _Resnet:
    def __init__(self, sd, p, dev, temb_off=None):
        g = lambda k: sd[p + k]
        self.g1, self.b1 = g("norm1.weight").to(dev, f32).contiguous(), g("norm1.bias").to(dev, f32).contiguous()
        self.g2, self.b2 = g("norm2.weight").to(dev, f32).contiguous(), g("norm2.bias").to(dev, f32).contiguous()
        self.w1 = pack_conv(g("conv1.weight")).to(dev, bf16)
        self.cb1 = g("conv1.bias").to(dev, f32).contiguous()
        w2 = pack_conv(g("conv2.weight"))
        cb2 = g("conv2.bias").to(f32)
        self.shortcut = (p + "conv_shortcut.weight") in sd
        if self.shortcut:                      # fold the 1x1 shortcut into conv2's K dimension
            w2 = torch.cat([w2, g("conv_shortcut.weight").flatten(1)], dim=1)
            cb2 = cb2 + g("conv_shortcut.bias").to(f32)
        self.w2 = w2.contiguous().to(dev, bf16)
        self.cb2 = cb2.to(dev).contiguous()
        self.cout = self.w1.shape[0]
        self.temb_off = temb_off


class _Transformer:
    def __init__(self, sd, p, dev, heads):
        g = lambda k: sd[p + k]
        t = "transformer_blocks.0."
        f = lambda k: g(k).to(dev, f32).contiguous()
        self.heads = heads
        self.gn_g, self.gn_b = f("norm.weight"), f("norm.bias")
        self.w_in, self.b_in = g("proj_in.weight").flatten(1).to(dev, bf16).contiguous(), f("proj_in.bias")
        self.w_out, self.b_out = g("proj_out.weight").flatten(1).to(dev, bf16).contiguous(), f("proj_out.bias")
        self.ln = [(f(t + f"norm{i}.weight"), f(t + f"norm{i}.bias")) for i in (1, 2, 3)]
        self.w_qkv = torch.cat([g(t + "attn1.to_q.weight"), g(t + "attn1.to_k.weight"), g(t + "attn1.to_v.weight")],
                               dim=0).to(dev, bf16).contiguous()
        self.w_o1, self.b_o1 = g(t + "attn1.to_out.0.weight").to(dev, bf16).contiguous(), f(t + "attn1.to_out.0.bias")
        self.w_q2 = g(t + "attn2.to_q.weight").to(dev, bf16).contiguous()
        self.w_kv2 = torch.cat([g(t + "attn2.to_k.weight"), g(t + "attn2.to_v.weight")], dim=0).to(dev, bf16).contiguous()
        self.w_o2, self.b_o2 = g(t + "attn2.to_out.0.weight").to(dev, bf16).contiguous(), f(t + "attn2.to_out.0.bias")
        wg, bg = interleave_geglu(g(t + "ff.net.0.proj.weight"), g(t + "ff.net.0.proj.bias"))
        self.w_gg, self.b_gg = wg.to(dev, bf16).contiguous(), bg.to(dev, f32).contiguous()
        self.w_ff, self.b_ff = g(t + "ff.net.2.weight").to(dev, bf16).contiguous(), f(t + "ff.net.2.bias")
        self.C = self.w_in.shape[0]
        self.kv = None            # cross-attention K/V of the current prompt batch: fp16 [Bu, 77, 2C]
        self.kv_bufs = {}         # persistent per batch size, so captured graphs keep valid addresses


class UNetB200:
    """``state_dict``: diffusers-keyed fp32/bf16 tensors (see weights.unet_param_shapes)."""

    def __init__(self, state_dict, in_channels: int = 4, device: str = "cuda", split_upsample: bool = True):
        shapes = unet_param_shapes(in_channels=in_channels)
        missing = [k for k in shapes if k not in state_dict]
        if missing:
            raise KeyError(f"UNet state dict is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, s in shapes.items():
            if tuple(state_dict[k].shape) != tuple(s):
                raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != {s}")
        # raw tensors go to the device first: the repacking below (permutes, concatenations, parity sums, casts) then runs
        # there instead of on the host -- seconds of the pipeline's cold start for 860 M parameters
        dev = device
        sd = {k: state_dict[k].to(dev, non_blocking=True) for k in shapes}
        self.device, self.in_channels = dev, in_channels
        self.boc, self.heads = (320, 640, 1280, 1280), 8
        self.split_upsample = split_upsample
        f = lambda k: sd[k].to(dev, f32).contiguous()
        # conv_in through im2col (K = 9*Cin padded to a multiple of 64)
        self.kpad_in = 64 * math.ceil(9 * in_channels / 64)
        w_in = torch.zeros((320, self.kpad_in), dtype=f32, device=dev)
        w_in[:, :9 * in_channels] = pack_conv(sd["conv_in.weight"].to(f32))
        self.w_conv_in, self.b_conv_in = w_in.to(dev, bf16).contiguous(), f("conv_in.bias")
        self.w_t1, self.b_t1 = sd["time_embedding.linear_1.weight"].to(dev, bf16).contiguous(), f("time_embedding.linear_1.bias")
        self.w_t2, self.b_t2 = sd["time_embedding.linear_2.weight"].to(dev, bf16).contiguous(), f("time_embedding.linear_2.bias")

        temb_w, temb_b, off = [], [], [0]

        def resnet(p):
            r = _Resnet(sd, p, dev, temb_off=off[0])
            temb_w.append(sd[p + "time_emb_proj.weight"]); temb_b.append(sd[p + "time_emb_proj.bias"])
            off[0] += r.cout
            return r

        self.down = []
        for i in range(4):
            blk = {"res": [], "attn": [], "down": None}
            for j in range(2):
                blk["res"].append(resnet(f"down_blocks.{i}.resnets.{j}."))
                if i < 3:
                    blk["attn"].append(_Transformer(sd, f"down_blocks.{i}.attentions.{j}.", dev, self.heads))
            if i < 3:
                blk["down"] = (pack_conv(sd[f"down_blocks.{i}.downsamplers.0.conv.weight"]).to(dev, bf16),
                               f(f"down_blocks.{i}.downsamplers.0.conv.bias"))
            self.down.append(blk)
        self.mid = {"res": [resnet("mid_block.resnets.0."), resnet("mid_block.resnets.1.")],
                    "attn": _Transformer(sd, "mid_block.attentions.0.", dev, self.heads)}
        self.up = []
        for i in range(4):
            blk = {"res": [], "attn": [], "up": None}
            for j in range(3):
                blk["res"].append(resnet(f"up_blocks.{i}.resnets.{j}."))
                if i > 0:
                    blk["attn"].append(_Transformer(sd, f"up_blocks.{i}.attentions.{j}.", dev, self.heads))
            if i < 3:
                w = sd[f"up_blocks.{i}.upsamplers.0.conv.weight"]
                bias = f(f"up_blocks.{i}.upsamplers.0.conv.bias")
                par = {(py, px): wp for py, px, wp in upsample_parity_weights(w)}
                blk["up"] = (torch.cat([par[(py, px)] for py in (0, 1) for px in (0, 1)], dim=0).to(dev, bf16).contiguous(),
                             pack_conv(w).to(dev, bf16), bias)
            self.up.append(blk)
        self.w_temb = torch.cat(temb_w, dim=0).to(dev, bf16).contiguous()
        self.b_temb = torch.cat(temb_b, dim=0).to(dev, f32).contiguous()
        self.g_out, self.b_out = f("conv_norm_out.weight"), f("conv_norm_out.bias")
        self.w_conv_out, self.b_conv_out = pack_conv(sd["conv_out.weight"]).to(dev, bf16), f("conv_out.bias")
        self.transformers = ([t for b in self.down for t in b["attn"]] + [self.mid["attn"]]
                             + [t for b in self.up for t in b["attn"]])

    # ------------------------------------------------------------------------------------------ prompt side
    def prepare_context(self, ctx: torch.Tensor):
        """ctx: [Bu, 77, 768] (fp32 or bf16) encoder_hidden_states.  Computes K/V of all 16 cross-attentions."""
        Bu, T, D = ctx.shape
        c = ctx.to(self.device)
        c = (ops.cast_bf16(c.to(f32).contiguous()) if c.dtype != bf16 else c.contiguous()).view(Bu * T, D)
        for t in self.transformers:
            buf = t.kv_bufs.get((Bu, T))
            if buf is None:
                buf = t.kv_bufs[(Bu, T)] = torch.empty((Bu, T, 2 * t.C), dtype=f16, device=self.device)
            ops.linear(c, t.w_kv2, images=Bu, out_bf16=buf.view(Bu, 1, T, 2 * t.C))
            t.kv = buf
        self.ctx_batch = Bu

    # ------------------------------------------------------------------------------------------ blocks
    def _resnet(self, r: _Resnet, x, skip, temb_all, want_bf16=False):
        N = x.shape[0]
        y1, raw = ops.groupnorm(x, r.g1, r.b1, eps=1e-5, silu=True, x2=skip, want_raw=r.shortcut)
        bn = temb_all[:, r.temb_off:r.temb_off + r.cout]
        # few output pixels (the 8x8 level): an fp32 output lets the GEMM split K across two clusters per tile
        small = x.shape[1] * x.shape[2] <= 64          # per image, never by batch size: results stay batch-invariant
        hb, hf = ops.conv2d(y1, r.w1, kh=3, kw=3, pad_t=1, pad_l=1, bias=r.cb1, bias_n=bn, out_bf16=not small,
                            out_f32=small)
        y2, _ = ops.groupnorm(hf if small else hb, r.g2, r.b2, eps=1e-5, silu=True)
        if r.shortcut:
            ob, of = ops.conv2d(y2, r.w2, kh=3, kw=3, pad_t=1, pad_l=1, x2=raw, bias=r.cb2, out_f32=True,
                                out_bf16=want_bf16)
        else:
            assert skip is None
            ob, of = ops.conv2d(y2, r.w2, kh=3, kw=3, pad_t=1, pad_l=1, bias=r.cb2, res=x, out_f32=True,
                                out_bf16=want_bf16)
        return of, ob

    def _transformer(self, t: _Transformer, x, want_bf16=False):
        N, H, W, Cc = x.shape
        heads, d = t.heads, Cc // t.heads
        M = N * H * W
        y, _ = ops.groupnorm(x, t.gn_g, t.gn_b, eps=1e-6, silu=False)
        _, tok = ops.conv2d(y, t.w_in, bias=t.b_in, out_f32=True)
        tok = tok.view(M, Cc)
        # self-attention
        l1 = ops.layernorm(tok, *t.ln[0])
        qkv, _ = ops.linear(l1, t.w_qkv, images=N, out_bf16=True, out_half=f16)         # attention operands in fp16 (like the reference)
        qkv = qkv.view(N, H * W, 3, heads, d)
        o = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
        ops.linear(o.view(M, Cc), t.w_o1, images=N, bias=t.b_o1, res=tok, out_f32=tok.view(N, 1, H * W, Cc))
        # cross-attention against the cached prompt K/V
        l2 = ops.layernorm(tok, *t.ln[1])
        q2, _ = ops.linear(l2, t.w_q2, images=N, out_bf16=True, out_half=f16)
        kv = t.kv.view(N, -1, 2, heads, d)
        o2 = ops.attention(q2.view(N, H * W, heads, d), kv[:, :, 0], kv[:, :, 1], d ** -0.5)
        ops.linear(o2.view(M, Cc), t.w_o2, images=N, bias=t.b_o2, res=tok, out_f32=tok.view(N, 1, H * W, Cc))
        # GEGLU feed-forward
        l3 = ops.layernorm(tok, *t.ln[2])
        gg, _ = ops.linear(l3, t.w_gg, images=N, bias=t.b_gg, act=RG_ACT_GEGLU, out_bf16=True)
        tb, _ = ops.linear(gg, t.w_ff, images=N, bias=t.b_ff, res=tok, out_bf16=True)
        ob, of = ops.conv2d(tb.view(N, H, W, Cc), t.w_out, bias=t.b_out, res=x, out_f32=True, out_bf16=want_bf16)
        return of, ob

    def _upsample_conv(self, up, xb, OH, OW):
        N, H, W, Cc = xb.shape
        wts, w_plain, bias = up
        if not self.split_upsample or (OH, OW) != (2 * H, 2 * W):
            # diffusers passes the skip's size to F.interpolate when the input is not divisible by 8
            u = ops.upsample_nearest(xb, OH, OW)
            _, of = ops.conv2d(u, w_plain, kh=3, kw=3, pad_t=1, pad_l=1, bias=bias, out_f32=True)
            return of
        out = torch.empty((N, 2 * H, 2 * W, Cc), dtype=f32, device=xb.device)
        # parity (py,px): output pixels (2j+py, 2i+px) = 2x2 conv over rows {j-1+py, j+py}, cols {i-1+px, i+px}; the four
        # kernels are stacked along Cout and run as ONE launch (rg_conv_t::parities = 4): at small batch each parity alone
        # fills a fraction of the SMs and costs a whole launch of latency
        ops.conv2d(xb, wts, kh=2, kw=2, OH=H, OW=W, bias=bias, out_f32=out, parities=4)
        return out

    # ------------------------------------------------------------------------------------------ forward
    def time_embedding(self, timesteps: torch.Tensor) -> torch.Tensor:
        """timesteps f32 [M] -> f32 [M, sum of resnet widths]: time_embedding MLP and all 22 ``time_emb_proj(silu(.))`` in
        three GEMMs.  Depends on the timestep only, so a sampling run computes it for ALL its steps in one go before the
        loop (rows are independent: the values equal those of a per-step evaluation bit for bit)."""
        M = timesteps.shape[0]
        te = ops.timestep_embedding(timesteps, 320)
        e1, _ = ops.linear(te, self.w_t1, images=M, bias=self.b_t1, act=RG_ACT_SILU, out_bf16=True)
        e2, _ = ops.linear(e1, self.w_t2, images=M, bias=self.b_t2, act=RG_ACT_SILU, out_bf16=True)     # silu(temb)
        _, temb_all = ops.linear(e2, self.w_temb, images=M, bias=self.b_temb, out_f32=True)
        return temb_all

    def forward(self, latents: torch.Tensor, timesteps: torch.Tensor | None = None,
                temb_all: torch.Tensor | None = None) -> torch.Tensor:
        """latents: f32 channels-last [n_mod, h, w, in_channels] (n_mod divides the context batch: under CFG the same
        latents feed both halves); timesteps: f32 [Bu], or ``temb_all``: the precomputed ``time_embedding`` rows as a
        [Bu, total] view (row stride 0 when every sample shares the timestep).  Returns eps f32 [Bu, h, w, 4]."""
        Bu = self.ctx_batch
        n_mod, h, w, cin = latents.shape
        assert cin == self.in_channels and Bu % n_mod == 0
        if temb_all is None:
            assert timesteps is not None and timesteps.shape[0] == Bu
            temb_all = self.time_embedding(timesteps)
        assert temb_all.shape[0] == Bu and temb_all.stride(1) == 1

        cols = ops.im2col_small(latents, Bu, 3, 1, 1, h, w, self.kpad_in)
        _, x = ops.conv2d(cols, self.w_conv_in, bias=self.b_conv_in, out_f32=True)
        skips = [x]
        xb = None
        for i, blk in enumerate(self.down):
            for j, r in enumerate(blk["res"]):
                last = j == len(blk["res"]) - 1 and blk["down"] is not None
                has_attn = bool(blk["attn"])
                x, xb = self._resnet(r, x, None, temb_all, want_bf16=last and not has_attn)
                if has_attn:
                    x, xb = self._transformer(blk["attn"][j], x, want_bf16=last)
                skips.append(x)
            if blk["down"] is not None:
                wd, bd = blk["down"]
                H, W = x.shape[1], x.shape[2]
                _, x = ops.conv2d(xb, wd, kh=3, kw=3, stride=2, pad_t=1, pad_l=1, OH=(H + 1) // 2, OW=(W + 1) // 2,
                                  bias=bd, out_f32=True)
                skips.append(x)
        x, _ = self._resnet(self.mid["res"][0], x, None, temb_all)
        x, _ = self._transformer(self.mid["attn"], x)
        x, _ = self._resnet(self.mid["res"][1], x, None, temb_all)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["res"]):
                last = j == len(blk["res"]) - 1 and blk["up"] is not None
                has_attn = bool(blk["attn"])
                x, xb = self._resnet(r, x, skips.pop(), temb_all, want_bf16=last and not has_attn)
                if has_attn:
                    x, xb = self._transformer(blk["attn"][j], x, want_bf16=last)
            if blk["up"] is not None:
                if h % 8 or w % 8:                # forward_upsample_size: target = next skip's spatial size
                    OH, OW = skips[-1].shape[1], skips[-1].shape[2]
                else:
                    OH, OW = 2 * x.shape[1], 2 * x.shape[2]
                x = self._upsample_conv(blk["up"], xb, OH, OW)
        y, _ = ops.groupnorm(x, self.g_out, self.b_out, eps=1e-5, silu=True)
        _, eps = ops.conv2d(y, self.w_conv_out, kh=3, kw=3, pad_t=1, pad_l=1, bias=self.b_conv_out, out_f32=True)
        return eps
