"""LPIPS (AlexNet, v0.1) on the sm_100a kernels -- the third per-image metric of the reference's evaluator.

Replaces ``lpips.LPIPS(net='alex')`` as the reference uses it (``/root/reference/src/metrics.py:67`` construction,
``:48-54`` ``preprocess_for_lpips`` -> [-1, 1] NCHW, ``:97-111`` ``calculate_lpips``: resize pred to gt, forward, ``.item()``).
Graph (lpips 0.1.x, ``net='alex'``, ``lpips=True``, ``spatial=False``): scaling layer ``(x - shift) / scale`` -> the five
ReLU taps of torchvision's AlexNet ``features`` (conv 11x11 s4 p2, maxpool, conv 5x5 p2, maxpool, 3 x conv 3x3) -> per tap:
unit-normalise along C (``x / (||x||_2 + 1e-10)``), squared difference, non-negative 1x1 "lin" head, spatial mean -> sum of
the five levels.

B200 mapping: every convolution is one ``rg_conv2d`` launch (tcgen05 implicit GEMM, bias + ReLU in the epilogue; the
11x11 / 5x5 kernels go through ``rg_im2col_small`` because the implicit-GEMM loader covers up to 3x3 taps), pooling is
``rg_maxpool3x3s2``, and ``rg_lpips_layer`` fuses normalise / diff / head / spatial sum per level.  Pred and gt run as one
batch of 2N.  Per-image values are assembled on the host in float64 from per-block partials in block order, so a value
does not depend on the batch (or the rank) an image was scored in.

Weights: the pretrained AlexNet + lin heads ship inside the ``lpips`` / torchvision wheels' download caches and are NOT
available offline; ``random_lpips_state_dict`` gives a seeded random-init network of the same architecture (values are
therefore unpinned -- what is checkable is the graph against the fp32 restatement ``oracle/lpips.py`` and the
bookkeeping).  ``load_lpips_state_dict`` accepts the real files when they exist.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import ops
from ._lib import RG_ACT_RELU
from .weights import pack_conv

bf16, f32 = ops.OPERAND_DTYPE, torch.float32      # bf16 = the build's 16-bit operand dtype (ops.py)
SHIFT = (-0.030, -0.088, -0.188)           # lpips.ScalingLayer
SCALE = (0.458, 0.448, 0.450)
CHANNELS = (64, 192, 384, 256, 256)
CONVS = (("features.0", 3, 64, 11), ("features.3", 64, 192, 5), ("features.6", 192, 384, 3), ("features.8", 384, 256, 3),
         ("features.10", 256, 256, 3))      # torchvision.models.alexnet().features indices


def lpips_param_shapes() -> "OrderedDict[str, tuple]":
    sh: OrderedDict = OrderedDict()
    for name, cin, cout, k in CONVS:
        sh[name + ".weight"] = (cout, cin, k, k)
        sh[name + ".bias"] = (cout,)
    for i, c in enumerate(CHANNELS):
        sh[f"lin{i}.model.1.weight"] = (1, c, 1, 1)
    return sh


def random_lpips_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random-init LPIPS-alex: He-normal convolutions, small biases, non-negative lin heads (0.1 |N(0,1)|, the
    constraint lpips trains under); every value bf16-representable so the fp32 checker and the kernels share weights."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: OrderedDict = OrderedDict()
    for name, shape in lpips_param_shapes().items():
        if name.startswith("lin"):
            t = torch.randn(shape, generator=g).abs() * 0.1
        elif name.endswith(".weight"):
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / math.prod(shape[1:]))
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        sd[name] = t.to(bf16).to(f32)
    return sd


def load_lpips_state_dict(alexnet_sd: dict, lin_sd: dict) -> "OrderedDict[str, torch.Tensor]":
    """torchvision ``alexnet`` state dict (``features.N.*``) + lpips' ``weights/v0.1/alex.pth`` (``linK.model.1.weight``)."""
    sd: OrderedDict = OrderedDict()
    for name, shape in lpips_param_shapes().items():
        src = lin_sd if name.startswith("lin") else alexnet_sd
        if name not in src:
            raise KeyError(f"LPIPS weights: missing {name}")
        if tuple(src[name].shape) != tuple(shape):
            raise ValueError(f"LPIPS weights: {name} has shape {tuple(src[name].shape)}, expected {shape}")
        sd[name] = src[name].detach().to(f32)
    return sd


class LPIPSB200:
    """``LPIPSB200(sd)(pred_u8, gt_u8)`` -> list of per-image LPIPS values (Python floats).
    ``pred_u8`` / ``gt_u8``: torch.uint8 CUDA tensors [N,H,W,3] of equal shape (H, W >= 35 so every tap is non-empty)."""

    def __init__(self, state_dict: dict, device: str = "cuda"):
        for k, s in lpips_param_shapes().items():
            if k not in state_dict or tuple(state_dict[k].shape) != tuple(s):
                raise ValueError(f"LPIPS state dict: bad or missing {k}")
        dev = self.device = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("LPIPSB200 runs on CUDA (sm_100a) only")
        self.w, self.b = [], []
        for name, cin, cout, k in CONVS:
            w = pack_conv(state_dict[name + ".weight"].to(f32))
            kpad = (w.shape[1] + 63) // 64 * 64
            wp = torch.zeros((cout, kpad), dtype=f32)
            wp[:, :w.shape[1]] = w
            self.w.append(wp.to(dev, bf16).contiguous())
            self.b.append(state_dict[name + ".bias"].to(dev, f32).contiguous())
        self.lin = [state_dict[f"lin{i}.model.1.weight"].reshape(-1).to(dev, f32).contiguous() for i in range(5)]
        self.sc_w = torch.diag(1.0 / torch.tensor(SCALE, dtype=f32)).to(dev).contiguous()
        self.sc_b = (-torch.tensor(SHIFT, dtype=f32) / torch.tensor(SCALE, dtype=f32)).to(dev).contiguous()

    def parameters(self):
        yield from self.w
        yield from self.b
        yield from self.lin

    def features(self, u8: torch.Tensor) -> list[torch.Tensor]:
        """uint8 [N,H,W,3] -> the five post-ReLU taps, bf16 channels-last."""
        N, H, W, _ = u8.shape
        x = ops.pointwise_small(ops.preprocess_u8(u8.contiguous()), self.sc_w, self.sc_b)      # [-1,1], then the scaling layer
        oh, ow = (H + 4 - 11) // 4 + 1, (W + 4 - 11) // 4 + 1
        cols = ops.im2col_small(x, N, 11, 4, 2, oh, ow, self.w[0].shape[1])
        r1, _ = ops.conv2d(cols, self.w[0], bias=self.b[0], act=RG_ACT_RELU, out_bf16=True)
        p1 = ops.maxpool3x3s2(r1)
        cols = ops.im2col_small(p1, N, 5, 1, 2, p1.shape[1], p1.shape[2], self.w[1].shape[1])
        r2, _ = ops.conv2d(cols, self.w[1], bias=self.b[1], act=RG_ACT_RELU, out_bf16=True)
        p2 = ops.maxpool3x3s2(r2)
        r3, _ = ops.conv2d(p2, self.w[2], kh=3, kw=3, pad_t=1, pad_l=1, bias=self.b[2], act=RG_ACT_RELU, out_bf16=True)
        r4, _ = ops.conv2d(r3, self.w[3], kh=3, kw=3, pad_t=1, pad_l=1, bias=self.b[3], act=RG_ACT_RELU, out_bf16=True)
        r5, _ = ops.conv2d(r4, self.w[4], kh=3, kw=3, pad_t=1, pad_l=1, bias=self.b[4], act=RG_ACT_RELU, out_bf16=True)
        return [r1, r2, r3, r4, r5]

    @torch.no_grad()
    def __call__(self, pred_u8: torch.Tensor, gt_u8: torch.Tensor) -> list[float]:
        if pred_u8.shape != gt_u8.shape or pred_u8.dim() != 4 or pred_u8.shape[3] != 3:
            raise ValueError(f"LPIPS needs two uint8 [N,H,W,3] batches of equal shape, got {tuple(pred_u8.shape)} / {tuple(gt_u8.shape)}")
        if not (pred_u8.is_cuda and gt_u8.is_cuda) or pred_u8.dtype != torch.uint8 or gt_u8.dtype != torch.uint8:
            raise ValueError("LPIPS needs torch.uint8 CUDA tensors")
        N, H, W, _ = pred_u8.shape
        if H < 35 or W < 35:
            raise ValueError("LPIPS (AlexNet) needs images of at least 35 x 35 pixels")
        with torch.cuda.device(self.device):
            feats = self.features(torch.cat([pred_u8, gt_u8], dim=0))
            parts = []
            for f, lin in zip(feats, self.lin):
                parts.append((ops.lpips_layer(f[:N].contiguous(), f[N:].contiguous(), lin), f.shape[1] * f.shape[2]))
            host = [(p.cpu().numpy().astype(np.float64), hw) for p, hw in parts]
        out = []
        for n in range(N):
            total = 0.0
            for p, hw in host:                       # levels in order, blocks in order: a fixed float64 summation
                s = 0.0
                for v in p[n]:
                    s += float(v)
                total += s / hw
            out.append(float(total))
        return out
