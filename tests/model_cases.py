"""Model-level parity cases: the CUDA path (through the C ABI) against the fp32 oracle restatement on the same
seeded inputs and the same random-init weights (weights.random_state_dict: bf16-representable values, so the only
difference is arithmetic).  Gates from BASELINE.json's north_star: per-step UNet rel-L2 <= 1e-2, final image
PSNR >= 40 dB."""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from image_restoration_and_enhancement_b200 import ops
from image_restoration_and_enhancement_b200.pipelines import (StableDiffusionImg2ImgPipeline,
                                                               StableDiffusionInpaintPipeline)
from image_restoration_and_enhancement_b200.unet import UNetB200
from image_restoration_and_enhancement_b200.vae import VAEB200
from image_restoration_and_enhancement_b200.weights import random_state_dict, unet_param_shapes, vae_param_shapes

DEV = "cuda"
UNET_TOL = 1e-2
PSNR_MIN = 40.0
LATENT_TOL = 2e-2      # latents along the trajectory (errors of up to 23 bf16 UNet evaluations + the VAE encoder accumulate)


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return float("inf") if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


_cache: dict = {}


def oracle_unet(in_channels=4, seed=0, sd=None):
    key = ("ounet", in_channels, seed)
    if key not in _cache:
        from oracle.unet import UNet2DConditionModel, UNetConfig
        if sd is None:
            sd = random_state_dict(unet_param_shapes(in_channels=in_channels), seed)
        m = UNet2DConditionModel(UNetConfig(in_channels=in_channels))
        m.load_state_dict(sd, strict=True)
        _cache[key] = (m.to(DEV).eval(), sd)
    return _cache[key]


def oracle_vae(seed=1):
    key = ("ovae", seed)
    if key not in _cache:
        from oracle.vae import AutoencoderKL
        sd = random_state_dict(vae_param_shapes(), seed)
        m = AutoencoderKL()
        m.load_state_dict(sd, strict=True)
        _cache[key] = (m.to(DEV).eval(), sd)
    return _cache[key]


def synth_image(seed: int, H: int = 512, W: int = 512) -> np.ndarray:
    """Smooth seeded test image (low-pass random field + gradient), uint8 HWC."""
    rng = np.random.default_rng(seed)
    low = rng.standard_normal((H // 32 + 2, W // 32 + 2, 3))
    t = torch.from_numpy(low).permute(2, 0, 1)[None].float()
    up = torch.nn.functional.interpolate(t, size=(H, W), mode="bicubic", align_corners=False)[0].permute(1, 2, 0).numpy()
    yy, xx = np.mgrid[0:H, 0:W]
    img = 127 + 50 * up + 30 * np.sin(xx / 37.0)[..., None] + 20 * (yy / H)[..., None]
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


# ------------------------------------------------------------------------------------------------ UNet
@torch.no_grad()
def case_unet(in_channels=4, B=1, h=64, w=64, cfg=True, t=501.0, seed=0):
    """The CUDA path runs FIRST (so a capped launch trace of this process shows rg:: kernels, not the oracle's cuDNN
    ones), the fp32 oracle second, on the same seeded inputs and the same weights."""
    _setup()
    key = ("unet", in_channels, seed)
    if key not in _cache:
        sd = random_state_dict(unet_param_shapes(in_channels=in_channels), seed)
        _cache[key] = (UNetB200(sd, in_channels=in_channels, device=DEV), sd)
    um, sd = _cache[key]
    g = torch.Generator().manual_seed(100 + seed)
    Bu = 2 * B if cfg else B
    lat = torch.randn((B, in_channels, h, w), generator=g).to(DEV)
    ctx = torch.randn((Bu, 77, 768), generator=g).to(DEV)
    um.prepare_context(ctx)
    ts = torch.full((Bu,), t, dtype=torch.float32, device=DEV)
    eps = um.forward(ops.nchw_to_nhwc(lat.contiguous()), ts)
    torch.cuda.synchronize()
    out = ops.nhwc_to_nchw(eps)
    om, _ = oracle_unet(in_channels, seed, sd)
    x_ref = torch.cat([lat] * 2) if cfg else lat
    ref = om(x_ref, torch.tensor(t, device=DEV), ctx)                       # [Bu,4,h,w]
    return rel_l2(out, ref), UNET_TOL


# ------------------------------------------------------------------------------------------------ VAE
@torch.no_grad()
def case_vae_encode(B=1, H=512, W=512, seed=1):
    _setup()
    om, sd = oracle_vae(seed)
    if ("vae", seed) not in _cache:
        _cache[("vae", seed)] = VAEB200(sd, device=DEV)
    vm = _cache[("vae", seed)]
    img = np.stack([synth_image(7 + i, H, W) for i in range(B)])
    x = ops.preprocess_u8(torch.from_numpy(img).to(DEV))
    ref = om.quant_conv(om.encoder(x.permute(0, 3, 1, 2).contiguous()))     # [B,8,h,w]
    mom = vm.encode_moments(x)
    torch.cuda.synchronize()
    return rel_l2(mom.permute(0, 3, 1, 2), ref), 1.5e-2


@torch.no_grad()
def case_vae_decode(B=1, h=64, w=64, seed=1):
    _setup()
    om, sd = oracle_vae(seed)
    if ("vae", seed) not in _cache:
        _cache[("vae", seed)] = VAEB200(sd, device=DEV)
    vm = _cache[("vae", seed)]
    g = torch.Generator().manual_seed(55)
    z = (torch.randn((B, 4, h, w), generator=g) * 0.18215).to(DEV)
    ref = om.decode(z / 0.18215)
    img = vm.decode(ops.nchw_to_nhwc(z.contiguous()))
    torch.cuda.synchronize()
    ref_u8 = ((ref / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8).cpu().numpy()
    out_u8 = ops.postprocess_u8(img).cpu().numpy()
    p = psnr_u8(out_u8, ref_u8)
    return rel_l2(img.permute(0, 3, 1, 2), ref), 2e-2, p


# ------------------------------------------------------------------------------------------------ whole pipeline
@torch.no_grad()
def case_pipeline(task="denoise", H=512, W=512, B=1, seed=0, graph=True, steps=None, oracle_images=None, step_stride=1):
    """Full sampling run vs the oracle pipeline: per-step guided-eps rel-L2 (each step fed the CUDA path's own
    UNet input, so the figure isolates one UNet evaluation) and PSNR of the final uint8 image.
    ``steps`` overrides num_inference_steps (BASELINE config 4 runs 50); ``oracle_images`` = indices of the batch the
    oracle re-runs and the image / latent comparisons cover (default: all); ``step_stride``: every n-th UNet step is
    replayed through the oracle UNet (default: every step)."""
    _setup()
    from oracle.pipelines import OraclePipeline, Trace
    params = {"denoise": dict(steps=20, strength=0.5, g=5.0, kind="pndm", cin=4),
              "colorize": dict(steps=30, strength=0.75, g=7.5, kind="pndm", cin=4),
              "sr": dict(steps=20, strength=0.8, g=0.0, kind="pndm", cin=4),
              "inpaint": dict(steps=30, strength=0.6, g=5.0, kind="ddim", cin=9)}[task]
    cin = params["cin"]
    if steps is not None:
        params["steps"] = steps
    idx = list(range(B)) if oracle_images is None else list(oracle_images)
    ou, usd = oracle_unet(cin, seed)
    ov, vsd = oracle_vae(seed + 1)
    cls = StableDiffusionInpaintPipeline if cin == 9 else StableDiffusionImg2ImgPipeline
    key = ("pipe", task if cin == 9 else "img2img", seed)
    if key not in _cache:
        from image_restoration_and_enhancement_b200.pipelines import _Tokenizer, make_text_encoder
        from image_restoration_and_enhancement_b200.schedulers import SCHEDULERS
        sched = SCHEDULERS["DDIMScheduler" if cin == 9 else "PNDMScheduler"]()
        _cache[key] = cls(dict(usd), dict(vsd), sched, make_text_encoder(seed + 2), _Tokenizer(None)).to(DEV)
    pipe = _cache[key]
    pipe.use_cuda_graph = graph
    g = torch.Generator().manual_seed(200 + seed)
    pe = torch.randn((1, 77, 768), generator=g).to(DEV)
    ne = torch.randn((1, 77, 768), generator=g).to(DEV)
    img = np.stack([synth_image(11 + i, H, W) for i in range(B)])
    gens = [torch.Generator(device=DEV).manual_seed(42) for _ in range(B)]
    trace: dict = {}
    kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, strength=params["strength"],
              num_inference_steps=params["steps"], guidance_scale=params["g"], generator=gens,
              output_type="np_u8", trace=trace)
    mask = None
    if cin == 9:
        mask = np.zeros((B, H, W), dtype=np.uint8)
        mask[:, H // 4:H // 2, W // 3:2 * W // 3] = 255
        out = pipe(image=img, mask_image=mask, **kw).images
    else:
        out = pipe(image=img, **kw).images
    torch.cuda.synchronize()

    # ---- oracle run on the same inputs and draws (B independent generators seeded 42 => identical draws per image)
    op = OraclePipeline(ou, ov, params["kind"])
    x = (torch.from_numpy(img).to(DEV).float() / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0
    otrace = Trace()
    refs = []
    for i in idx:
        gi = torch.Generator(device=DEV).manual_seed(42)
        if cin == 9:
            m = (torch.from_numpy(mask[i:i + 1]).to(DEV).float() / 255.0 >= 0.5).float()[:, None]
            refs.append(op.inpaint(x[i:i + 1], m, pe, ne, strength=params["strength"],
                                   num_inference_steps=params["steps"], guidance_scale=params["g"], generator=gi,
                                   trace=otrace if i == idx[0] else None))
        else:
            refs.append(op.img2img(x[i:i + 1], pe, ne, strength=params["strength"],
                                   num_inference_steps=params["steps"], guidance_scale=params["g"], generator=gi,
                                   trace=otrace if i == idx[0] else None))
    ref = np.concatenate(refs)
    i0 = idx[0]                                          # the image whose latent trajectory the oracle traced
    res = {"psnr": min(psnr_u8(out[i:i + 1], ref[k:k + 1]) for k, i in enumerate(idx)), "steps": len(trace["timesteps"]),
           "timesteps_match": trace["timesteps"] == otrace.timesteps}
    res["init_latents_rel"] = rel_l2(ops.nhwc_to_nchw(trace["init_latents"])[i0:i0 + 1], otrace.init_latents)
    res["final_latents_rel"] = rel_l2(ops.nhwc_to_nchw(trace["final_latents"])[i0:i0 + 1], otrace.final_latents)
    # per-step UNet error with the CUDA path's own inputs replayed through the oracle UNet
    do_cfg = params["g"] > 1.0
    Bu = 2 * B if do_cfg else B
    embeds = torch.cat([ne.repeat(B, 1, 1), pe.repeat(B, 1, 1)]) if do_cfg else pe.repeat(B, 1, 1)
    step_err = []
    for i in range(0, len(trace["timesteps"]), step_stride):      # EVERY step of the run unless a stride is given
        xin = ops.nhwc_to_nchw(trace["unet_in"][i].contiguous())
        xin = torch.cat([xin] * 2) if do_cfg else xin
        ref_eps = ou(xin, torch.tensor(float(trace["timesteps"][i]), device=DEV), embeds)
        got = ops.nhwc_to_nchw(trace["eps_uc"][i].contiguous())
        step_err.append(rel_l2(got, ref_eps))
    res["unet_step_rel"] = step_err
    # scheduler-state parity along the whole trajectory: the CUDA path's latents after every step against the oracle's
    res["latents_rel_per_step"] = [rel_l2(ops.nhwc_to_nchw(a.contiguous())[i0:i0 + 1], b)
                                   for a, b in zip(trace["latents"], otrace.latents)]
    return res


@torch.no_grad()
def case_lpips(seed=3):
    """LPIPS-alex on the sm_100a kernels (lpips.LPIPSB200) against the fp32 restatement oracle/lpips.py on the same seeded
    random-init weights: 512x512 pairs of three kinds (noisy, blurred, unrelated) and an odd size.  Also: the value of an
    image does not depend on the batch, and MetricsCalculator(device="cuda", use_lpips=True) returns the same number.
    Returns (worst relative error, tolerance)."""
    _setup()
    import cv2
    from oracle.lpips import LPIPSAlex, preprocess_for_lpips
    from image_restoration_and_enhancement_b200.lpips import LPIPSB200, random_lpips_state_dict
    from image_restoration_and_enhancement_b200.metrics import MetricsCalculator
    sd = random_lpips_state_dict(seed)
    mine = LPIPSB200(sd, device=DEV)
    ref_model = LPIPSAlex().load_lpips_state_dict(sd).to(DEV)
    rng = np.random.default_rng(seed)
    worst = 0.0
    for (H, W) in ((512, 512), (333, 500)):
        gt = np.stack([synth_image(20 + i, H, W) for i in range(3)])
        pred = gt.copy()
        pred[0] = np.clip(gt[0].astype(np.float32) + rng.normal(0, 12, gt[0].shape), 0, 255).astype(np.uint8)
        pred[1] = cv2.GaussianBlur(gt[1], (9, 9), 0)
        pred[2] = synth_image(99, H, W)
        got = mine(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV))
        want = [float(ref_model(preprocess_for_lpips(p).to(DEV), preprocess_for_lpips(g).to(DEV))) for p, g in zip(pred, gt)]
        alone = mine(torch.from_numpy(pred[1:2]).to(DEV), torch.from_numpy(gt[1:2]).to(DEV))
        assert alone[0] == got[1], "LPIPS of an image must not depend on the batch"
        for a, b in zip(got, want):
            # bf16 feature maps put a rounding-noise floor of ~1e-5 under the distance of near-identical images (the blurred
            # smooth pair scores 4e-5): relative error against max(|ref|, 1e-3)
            worst = max(worst, abs(a - b) / max(abs(b), 1e-3))
    calc = MetricsCalculator(use_lpips=True, device=DEV, lpips_seed=seed)
    m = calc.calculate_all(pred[0], gt[0])
    assert set(m) == {"psnr", "ssim", "lpips"} and m["lpips"] == got[0]
    return worst, 2e-2


# ------------------------------------------------------------------------------------------------ timing helper
@torch.no_grad()
def case_clip_text(seed=7):
    """CLIP text encoder on librestoragen (text_encoder.py) against transformers' CLIPTextModel in fp32 on the same
    weights and ids: the reference's default prompts and the empty negative prompt.  Returns (rel-L2, tolerance)."""
    _setup()
    import json
    from pathlib import Path
    import image_restoration_and_enhancement_b200 as pkg
    from image_restoration_and_enhancement_b200.pipelines import make_text_encoder
    from image_restoration_and_enhancement_b200.text_encoder import CLIPTextB200
    ref_model = make_text_encoder(seed).to(DEV)
    table = json.loads((Path(pkg.__file__).parent / "data" / "default_prompt_ids.json").read_text())
    ids = torch.tensor(list(table.values()), dtype=torch.long, device=DEV)
    with torch.no_grad():
        want = ref_model(ids)[0].float()
    mine = CLIPTextB200(ref_model.state_dict(), device=DEV)
    got = mine(ids)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (ids.shape[0], 77, 768) and got.dtype == torch.float32
    one = mine(ids[:1])                                   # batch-invariant: a prompt alone gives the same embedding
    assert torch.equal(one, got[:1])
    return rel_l2(got, want), UNET_TOL


def time_unet(B=8, cfg=True, h=64, w=64, iters=5, graph=True, in_channels=4):
    _setup()
    sd = random_state_dict(unet_param_shapes(in_channels=in_channels), 0)
    um = UNetB200(sd, in_channels=in_channels, device=DEV)
    Bu = 2 * B if cfg else B
    lat = torch.randn((B, h, w, in_channels), device=DEV)
    ctx = torch.randn((Bu, 77, 768), device=DEV)
    um.prepare_context(ctx)
    ts = torch.full((Bu,), 500.0, device=DEV)
    n0 = ops.launch_count()
    um.forward(lat, ts)
    launches = ops.launch_count() - n0
    torch.cuda.synchronize()
    if graph:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            um.forward(lat, ts)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            um.forward(lat, ts)
        fn = gr.replay
    else:
        fn = lambda: um.forward(lat, ts)
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 0.8033e12 * Bu * (h * w) / 4096.0
    return {"B": B, "Bu": Bu, "ms": ms, "tflops": flops / ms / 1e9, "launches": launches, "graph": graph}
