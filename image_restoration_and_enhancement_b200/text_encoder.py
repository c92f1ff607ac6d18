"""CLIPTextModel forward pass on the sm_100a kernels (SURVEY 8f "f3").

The reference's pipelines call ``text_encoder(ids)[0]`` once per prompt (diffusers ``encode_prompt``; architecture in
``outputs/models/denoising/best/text_encoder/config.json``: 12 pre-LN layers, 768 wide, 12 heads of 64, MLP 3072 with
quick_gelu, 77 learned positions, causal mask, final LayerNorm).  Here every layer is

    LayerNorm (rg_layernorm) -> q|k|v in one GEMM (rg_conv2d as linear, bias fused) -> causal attention on tcgen05
    (rg_attention, causal = 1) -> out_proj GEMM with bias + residual fused, in place on the fp32 stream
    -> LayerNorm -> fc1 GEMM (+bias) -> quick_gelu (rg_quick_gelu_bf16) -> fc2 GEMM with bias + residual fused

on a fp32 residual stream with bf16 GEMM operands, exactly like the UNet's transformer blocks.  Weights come from a
transformers-layout state dict (``text_model.encoder.layers.N.self_attn.q_proj.weight`` ...), so a real checkpoint's
``text_encoder/model.safetensors`` loads unchanged.  Nothing here computes with torch.
"""
from __future__ import annotations

import torch

from . import ops

bf16, f32 = ops.OPERAND_DTYPE, torch.float32      # bf16 = the build's 16-bit operand dtype (ops.py)


class CLIPTextB200:
    def __init__(self, sd: dict, device: str = "cuda", heads: int = 12, eps: float = 1e-5):
        dev = torch.device(device)
        pre = "text_model." if any(k.startswith("text_model.") for k in sd) else ""
        g = lambda k: sd[pre + k]
        f = lambda k: g(k).to(dev, f32).contiguous()
        w = lambda k: g(k).to(dev, bf16).contiguous()
        self.device, self.heads, self.eps = dev, heads, eps
        self.tok = f("embeddings.token_embedding.weight")
        self.pos = f("embeddings.position_embedding.weight")
        self.vocab, self.C = self.tok.shape
        self.T = self.pos.shape[0]
        self.layers = []
        i = 0
        while f"{pre}encoder.layers.{i}.layer_norm1.weight" in sd:
            p = f"encoder.layers.{i}."
            self.layers.append(dict(
                ln1=(f(p + "layer_norm1.weight"), f(p + "layer_norm1.bias")),
                w_qkv=torch.cat([g(p + f"self_attn.{n}_proj.weight") for n in "qkv"], dim=0).to(dev, bf16).contiguous(),
                b_qkv=torch.cat([g(p + f"self_attn.{n}_proj.bias") for n in "qkv"], dim=0).to(dev, f32).contiguous(),
                w_o=w(p + "self_attn.out_proj.weight"), b_o=f(p + "self_attn.out_proj.bias"),
                ln2=(f(p + "layer_norm2.weight"), f(p + "layer_norm2.bias")),
                w_fc1=w(p + "mlp.fc1.weight"), b_fc1=f(p + "mlp.fc1.bias"),
                w_fc2=w(p + "mlp.fc2.weight"), b_fc2=f(p + "mlp.fc2.bias")))
            i += 1
        if not self.layers:
            raise OSError("state dict holds no CLIP text encoder layers")
        self.ln_f = (f("final_layer_norm.weight"), f("final_layer_norm.bias"))

    def parameters(self):
        yield self.tok
        yield self.pos
        for L in self.layers:
            for v in L.values():
                if isinstance(v, tuple):
                    yield from v
                else:
                    yield v
        yield from self.ln_f

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        """ids int [B, T<=77] -> last_hidden_state fp32 [B, T, 768] (after final_layer_norm, as ``text_encoder(ids)[0]``)."""
        if ids.dim() != 2 or ids.shape[1] > self.T:
            raise ValueError(f"ids must be [B, T<={self.T}], got {tuple(ids.shape)}")
        B, T = ids.shape
        C, H = self.C, self.heads
        d = C // H
        M = B * T
        with torch.cuda.device(self.device):
            x = ops.embed_tokens(ids.to(self.device, torch.int32).contiguous(), self.tok, self.pos)      # fp32 [M, C]
            for L in self.layers:
                h = ops.layernorm(x, *L["ln1"], eps=self.eps)
                qkv, _ = ops.linear(h, L["w_qkv"], bias=L["b_qkv"], out_bf16=True)
                qkv = qkv.view(B, T, 3, H, d)
                o = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5, causal=True)
                ops.linear(o.view(M, C), L["w_o"], bias=L["b_o"], res=x, out_f32=x.view(1, 1, M, C))
                h = ops.layernorm(x, *L["ln2"], eps=self.eps)
                u, _ = ops.linear(h, L["w_fc1"], bias=L["b_fc1"], out_bf16=True)
                ops.quick_gelu_(u)
                ops.linear(u, L["w_fc2"], bias=L["b_fc2"], res=x, out_f32=x.view(1, 1, M, C))
            y = ops.layernorm(x, *self.ln_f, eps=self.eps)
            return ops.cast_bf16_f32(y).view(B, T, C)

    __call__ = forward
