"""ctypes binding of librestoragen.so (declared in include/restoragen.h).

The library is the product's only compute path: if it cannot be loaded this module raises --
there is no CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_HERE = Path(__file__).resolve().parent
# RESTORAGEN_OPERAND_DTYPE=fp16 selects the fp16 parity build (same sources, -DRG_OPERAND_F16): every 16-bit activation and
# weight is fp16 and the tensor cores multiply fp16 operands, like the reference on CUDA (src/inference.py:57).  The
# choice is per process and made at import; the host modules allocate their 16-bit tensors as OPERAND_DTYPE_NAME.
OPERAND_DTYPE_NAME = {"": "bfloat16", "bf16": "bfloat16", "bfloat16": "bfloat16", "fp16": "float16", "f16": "float16",
                      "float16": "float16"}.get(os.environ.get("RESTORAGEN_OPERAND_DTYPE", "").lower())
if OPERAND_DTYPE_NAME is None:
    raise ValueError("RESTORAGEN_OPERAND_DTYPE must be bf16 or fp16")
LIB_PATH = _HERE / ("librestoragen_f16.so" if OPERAND_DTYPE_NAME == "float16" else "librestoragen.so")

RG_ACT_NONE, RG_ACT_SILU, RG_ACT_GEGLU, RG_ACT_RELU = 0, 1, 2, 3
RG_DT_BF16, RG_DT_F32, RG_DT_F16 = 0, 1, 2


class RgAct(C.Structure):
    _fields_ = [("data", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("stride_n", C.c_int64), ("stride_h", C.c_int64), ("stride_w", C.c_int64)]


class RgConv(C.Structure):
    _fields_ = [("x", RgAct), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
                ("pad_t", C.c_int32), ("pad_l", C.c_int32), ("OH", C.c_int32), ("OW", C.c_int32),
                ("has_x2", C.c_int32), ("x2", RgAct), ("w", C.c_void_p), ("w_ld", C.c_int64), ("Cout", C.c_int32),
                ("bias", C.c_void_p), ("bias_n", C.c_void_p), ("bias_n_ld", C.c_int64), ("res", C.c_void_p), ("res_dtype", C.c_int32),
                ("out_bf16", C.c_void_p), ("out_f32", C.c_void_p),
                ("out_stride_n", C.c_int64), ("out_stride_h", C.c_int64), ("out_stride_w", C.c_int64),
                ("act", C.c_int32), ("scale", C.c_float), ("out16_dtype", C.c_int32),
                ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64), ("parities", C.c_int32)]


class RgAttn(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
                ("B", C.c_int32), ("heads", C.c_int32), ("d", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32),
                ("q_stride_b", C.c_int64), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
                ("k_stride_b", C.c_int64), ("k_stride_t", C.c_int64), ("k_stride_h", C.c_int64),
                ("v_stride_b", C.c_int64), ("v_stride_t", C.c_int64), ("v_stride_h", C.c_int64),
                ("o_stride_b", C.c_int64), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
                ("scale", C.c_float), ("dtype", C.c_int32), ("causal", C.c_int32)]


class RgGn(C.Structure):
    _fields_ = [("x1", C.c_void_p), ("C1", C.c_int32), ("x2", C.c_void_p), ("C2", C.c_int32),
                ("in_dtype", C.c_int32), ("N", C.c_int32), ("HW", C.c_int64), ("groups", C.c_int32),
                ("eps", C.c_float), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("sums", C.c_void_p),
                ("y", C.c_void_p), ("raw", C.c_void_p), ("silu", C.c_int32)]


class RgSched(C.Structure):
    _fields_ = [("eps_uc", C.c_void_p), ("sample", C.c_void_p), ("ets", C.c_void_p), ("cur_sample", C.c_void_p),
                ("n", C.c_int64), ("do_cfg", C.c_int32), ("guidance", C.c_float), ("store_slot", C.c_int32),
                ("w", C.c_float * 5), ("use_cur", C.c_int32), ("save_cur", C.c_int32),
                ("c_sample", C.c_float), ("c_eps", C.c_float)]


# name -> (restype, argtypes); every symbol include/restoragen.h declares
_i32, _i64, _f32, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SIGNATURES = {
    "rg_last_error": (C.c_char_p, []),
    "rg_version": (C.c_int, []),
    "rg_operand_dtype": (C.c_int, []),
    "rg_launch_count": (C.c_int64, []),
    "rg_device_sm_count": (C.c_int, []),
    "rg_set_pdl": (C.c_int, [C.c_int]),
    "rg_conv2d": (C.c_int, [C.POINTER(RgConv), _p]),
    "rg_attention": (C.c_int, [C.POINTER(RgAttn), _p]),
    "rg_softmax_rows": (C.c_int, [_p, _i64, _i32, _i64, _p]),
    "rg_groupnorm_stats": (C.c_int, [C.POINTER(RgGn), _p]),
    "rg_groupnorm_apply": (C.c_int, [C.POINTER(RgGn), _p]),
    "rg_groupnorm": (C.c_int, [C.POINTER(RgGn), _p]),
    "rg_layernorm": (C.c_int, [_p, _i32, _i64, _i32, _p, _p, _f32, _p, _p]),
    "rg_timestep_embedding": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "rg_sched_step": (C.c_int, [C.POINTER(RgSched), _p]),
    "rg_im2col_small": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_upsample_nearest": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_nchw_to_nhwc": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_nhwc_to_nchw": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_preprocess_u8": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "rg_postprocess_u8": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_vae_sample": (C.c_int, [_p, _i64, _p, _p, _i64, _f32, _i32, _f32, _f32, _p, _p]),
    "rg_pack_unet_input": (C.c_int, [_p, _p, _p, _i64, _p, _p]),
    "rg_pointwise_small": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _f32, _p, _p]),
    "rg_mask_nearest": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_scale_f32": (C.c_int, [_p, _f32, _i64, _p, _p]),
    "rg_cast_f32_bf16": (C.c_int, [_p, _i64, _p, _p]),
    "rg_memset_zero": (C.c_int, [_p, _i64, _p]),
    "rg_embed_tokens": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_quick_gelu_bf16": (C.c_int, [_p, _i64, _p]),
    "rg_cast_bf16_f32": (C.c_int, [_p, _i64, _p, _p]),
    "rg_maxpool3x3s2": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "rg_lpips_layer_blocks": (C.c_int, [_i32]),
    "rg_lpips_layer": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p]),
    "rg_metrics_sse_u8": (C.c_int, [_p, _p, _i32, _i64, _p, _p]),
    "rg_metrics_ssim_chunks": (C.c_int, [_i32, _i32]),
    "rg_metrics_ssim_u8": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, C.c_double, C.c_double, C.c_double, _p, _p, _p]),
}


class RestoragenError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen librestoragen.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RestoragenError(
            f"{LIB_PATH} is missing: build it with `python -m image_restoration_and_enhancement_b200.build` "
            "(or __graft_entry__.build()).  There is no fallback path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    want = RG_DT_F16 if OPERAND_DTYPE_NAME == "float16" else RG_DT_BF16
    if lib.rg_operand_dtype() != want:
        raise RestoragenError(f"{LIB_PATH} was built for another 16-bit operand type than {OPERAND_DTYPE_NAME}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rg_last_error().decode(errors="replace")
        raise RestoragenError(f"{what or 'librestoragen'} failed (code {rc}): {msg}")
