"""Small target for `ncu --set full`: a few launches of the dominant kernels at UNet batch-16 shapes."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kernel_cases as kc
from image_restoration_and_enhancement_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "conv"
torch.manual_seed(0)
if which == "conv":          # resnet conv 320->320 3x3 at 64x64, 16 samples: M=65536, N=320, K=2880 (120.8 GFLOP)
    x = torch.randn((16, 64, 64, 320), device="cuda").to(torch.bfloat16)
    w = (torch.randn((320, 2880), device="cuda") / 53.0).to(torch.bfloat16)
    b = torch.randn((320,), device="cuda")
    for _ in range(4):
        ops.conv2d(x, w, kh=3, kw=3, pad_t=1, pad_l=1, bias=b, out_bf16=True)
elif which == "linear":      # attention out-projection with fp32 residual in/out: M=65536, N=320, K=320
    x = torch.randn((65536, 320), device="cuda").to(torch.bfloat16)
    w = (torch.randn((320, 320), device="cuda") / 18.0).to(torch.bfloat16)
    r = torch.randn((65536, 320), device="cuda")
    for _ in range(4):
        ops.linear(x, w, res=r, out_f32=r.view(1, 1, 65536, 320))
elif which == "qkv":         # q|k|v projection at 64x64 latents: M=65536, N=960, K=320, bf16 out (epilogue / HBM bound)
    x = torch.randn((65536, 320), device="cuda").to(torch.bfloat16)
    w = (torch.randn((960, 320), device="cuda") / 18.0).to(torch.bfloat16)
    for _ in range(4):
        ops.linear(x, w, out_bf16=True)
elif which == "geglu":       # ff.net.0 at 64x64 latents: M=65536, N=2560 (-> 1280 after GEGLU), K=320: the largest transformer GEMM
    from image_restoration_and_enhancement_b200._lib import RG_ACT_GEGLU
    x = torch.randn((65536, 320), device="cuda").to(torch.bfloat16)
    w = (torch.randn((2560, 320), device="cuda") / 18.0).to(torch.bfloat16)
    b = torch.randn((2560,), device="cuda")
    for _ in range(3):
        ops.linear(x, w, bias=b, act=RG_ACT_GEGLU, out_bf16=True)
elif which == "attn":        # self-attention at 64x64 latents: B=16, 8 heads, d=40, N=4096
    qkv = torch.randn((16, 4096, 3, 8, 40), device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 40 ** -0.5)
elif which == "attn16":      # the UNet's path: fp16 operands, 2 issuers, staggered warpgroups
    qkv = torch.randn((16, 4096, 3, 8, 40), device="cuda").half()
    for _ in range(3):
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 40 ** -0.5)
elif which == "gn":          # GroupNorm+SiLU: UNet 64x64x320 fp32 (stats + apply), VAE 512x512x128 bf16, UNet 16x16x1280 (one pass)
    g = torch.ones(1280, device="cuda"); bta = torch.zeros(1280, device="cuda")
    x = torch.randn((16, 64, 64, 320), device="cuda")
    xv = torch.randn((8, 512, 512, 128), device="cuda").to(torch.bfloat16)
    xs = torch.randn((16, 16, 16, 1280), device="cuda")
    for _ in range(2):
        ops.groupnorm(x, g[:320].contiguous(), bta[:320].contiguous(), silu=True)
        ops.groupnorm(xv, g[:128].contiguous(), bta[:128].contiguous(), eps=1e-6, silu=True)
        ops.groupnorm(xs, g, bta, silu=True)
elif which == "conv_b2":     # config 1 (UNet batch 2): resnet conv 320->320 3x3 at 64x64, M=8192, N=320, K=2880 (15.1 GFLOP)
    x = torch.randn((2, 64, 64, 320), device="cuda").to(torch.bfloat16)
    w = (torch.randn((320, 2880), device="cuda") / 53.0).to(torch.bfloat16)
    b = torch.randn((320,), device="cuda")
    for _ in range(4):
        ops.conv2d(x, w, kh=3, kw=3, pad_t=1, pad_l=1, bias=b, out_bf16=True)
elif which == "splitk8":     # config-1 shape: 8x8-level resnet conv at UNet batch 2, M=128, N=1280, K=11520 -> 8 K slices
    x = torch.randn((2, 8, 8, 1280), device="cuda").to(torch.bfloat16)
    w = (torch.randn((1280, 11520), device="cuda") / 107.0).to(torch.bfloat16)
    b = torch.randn((1280,), device="cuda")
    r = torch.randn((2, 8, 8, 1280), device="cuda")
    for _ in range(4):
        ops.conv2d(x, w, kh=3, kw=3, pad_t=1, pad_l=1, bias=b, res=r, out_f32=True)
elif which == "splitk4":     # 16x16-level conv at UNet batch 2: M=512, N=1280, K=11520 -> 4 K slices
    x = torch.randn((2, 16, 16, 1280), device="cuda").to(torch.bfloat16)
    w = (torch.randn((1280, 11520), device="cuda") / 107.0).to(torch.bfloat16)
    b = torch.randn((1280,), device="cuda")
    for _ in range(4):
        ops.conv2d(x, w, kh=3, kw=3, pad_t=1, pad_l=1, bias=b, out_bf16=True)
elif which == "ln":          # LayerNorm 65536 x 320 fp32 -> bf16
    x = torch.randn((65536, 320), device="cuda")
    g = torch.ones(320, device="cuda"); bta = torch.zeros(320, device="cuda")
    for _ in range(3):
        ops.layernorm(x, g, bta)
torch.cuda.synchronize()
print("done", which)
