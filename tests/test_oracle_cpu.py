"""CPU tests: the oracle against the reference's known answers / committed golden vectors, and the host logic."""
import json
import math
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


# ------------------------------------------------------------------------------------------------ known answers
def test_param_counts_match_reference_log():
    """859,520,964 is the UNet size the reference logged (outputs/models/colorization/training_colorization.log:30)."""
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from oracle.vae import AutoencoderKL
    with torch.device("meta"):
        u4, u9, v = UNet2DConditionModel(), UNet2DConditionModel(UNetConfig(in_channels=9)), AutoencoderKL()
    assert sum(p.numel() for p in u4.parameters()) == 859_520_964
    assert sum(p.numel() for p in u9.parameters()) == 859_535_364
    assert sum(p.numel() for p in v.parameters()) == 83_653_863


def test_product_shape_tables_match_oracle_modules():
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from oracle.vae import AutoencoderKL
    from image_restoration_and_enhancement_b200.weights import unet_param_shapes, vae_param_shapes
    with torch.device("meta"):
        for cin in (4, 9):
            sd = {k: tuple(v.shape) for k, v in UNet2DConditionModel(UNetConfig(in_channels=cin)).state_dict().items()}
            assert sd == dict(unet_param_shapes(in_channels=cin))
        sd = {k: tuple(v.shape) for k, v in AutoencoderKL().state_dict().items()}
        assert sd == dict(vae_param_shapes())


def test_scheduler_tables_golden():
    """Timestep lists of SURVEY.md 8(d) and the scaled-linear beta schedule end points."""
    from oracle.schedulers import DDIMScheduler, PNDMScheduler, get_timesteps
    gold = json.loads((GOLD / "scheduler_tables.json").read_text())
    s = PNDMScheduler(); s.set_timesteps(20)
    assert [int(t) for t in s.timesteps] == gold["denoise_pndm_20_0.5"]["full"]
    assert len(s.timesteps) == 21 and int(s.timesteps[1]) == int(s.timesteps[2]) == 901
    ts, n = get_timesteps(s, 20, 0.5)
    assert [int(t) for t in ts] == [501, 451, 401, 351, 301, 251, 201, 151, 101, 51, 1] and n == 10
    s.set_timesteps(30)
    assert [int(t) for t in get_timesteps(s, 30, 0.75)[0]] == gold["colorize_pndm_30_0.75"]["sliced"]
    assert len(gold["colorize_pndm_30_0.75"]["sliced"]) == 23
    s.set_timesteps(50)
    assert len(get_timesteps(s, 50, 0.8)[0]) == 41
    d = DDIMScheduler(); d.set_timesteps(30)
    assert [int(t) for t in d.timesteps][:3] == [958, 925, 892]
    assert [int(t) for t in get_timesteps(d, 30, 0.6)[0]] == gold["inpaint_ddim_30_0.6"]["sliced"]
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    assert abs(float(betas[0]) - 0.00085) < 1e-9 and abs(float(betas[-1]) - 0.012) < 1e-8
    assert abs(float(s.alphas_cumprod[0]) - gold["alphas_cumprod"]["0"]) < 1e-9
    assert abs(float(s.alphas_cumprod[999]) - gold["alphas_cumprod"]["999"]) < 1e-9


def test_prompt_token_ids_golden():
    """ids produced by the shipped tokenizer files (SURVEY.md 8c golden vector) are what the product falls back to."""
    from image_restoration_and_enhancement_b200.pipelines import _Tokenizer
    gold = json.loads((GOLD / "prompt_ids.json").read_text())
    tok = _Tokenizer(None)
    ids = tok("clean high quality photo, no noise, sharp details")
    assert ids[:13] == [49406, 3772, 1400, 3027, 1125, 267, 871, 9307, 267, 8157, 2353, 49407, 49407]
    assert tok("") == [49406] + [49407] * 76
    assert len(ids) == 77 and all(tok(p) == v for p, v in gold.items())
    with pytest.raises(OSError):
        tok("a prompt that is not built in")


def test_timestep_embedding_layout():
    from oracle.unet import timestep_embedding
    e = timestep_embedding(torch.tensor([0.0, 10.0]), 320, True, 0)
    assert e.shape == (2, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))   # cos | sin
    assert abs(float(e[1, 160]) - math.sin(10.0)) < 1e-6


# ------------------------------------------------------------------------------------------------ product scheduler == oracle scheduler
def _emulate(plans, eps_list, sample, g=None):
    """The arithmetic of rg_sched_step, in torch on the CPU."""
    ets = [None] * 4
    cur = None
    for p, e in zip(plans, eps_list):
        if p.store_slot >= 0:
            ets[p.store_slot] = e
        mix = p.w[4] * e
        for i in range(4):
            if p.w[i] != 0.0 and i != p.store_slot:
                mix = mix + p.w[i] * ets[i]
            elif p.w[i] != 0.0:
                mix = mix + p.w[i] * e
        base = cur if p.use_cur else sample
        if p.save_cur:
            cur = sample
        sample = p.c_sample * base - p.c_eps * mix
    return sample


@pytest.mark.parametrize("kind,steps,strength", [("pndm", 20, 0.5), ("pndm", 30, 0.75), ("pndm", 20, 0.8),
                                                 ("pndm", 50, 0.8), ("ddim", 30, 0.6), ("ddim", 30, 1.0)])
def test_step_plans_reproduce_oracle_scheduler(kind, steps, strength):
    from oracle import schedulers as osch
    from image_restoration_and_enhancement_b200 import schedulers as psch
    o = osch.PNDMScheduler() if kind == "pndm" else osch.DDIMScheduler()
    p = psch.PNDMScheduler() if kind == "pndm" else psch.DDIMScheduler()
    o.set_timesteps(steps); p.set_timesteps(steps)
    ots, _ = osch.get_timesteps(o, steps, strength)
    pts = p.get_timesteps(steps, strength)
    assert [int(t) for t in ots] == pts
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(4, 8, 8, generator=g, dtype=torch.float64)
    eps = [torch.randn(4, 8, 8, generator=g, dtype=torch.float64) for _ in pts]
    ref = x0.clone()
    for t, e in zip(ots, eps):
        ref = o.step(e, t, ref)
    got = _emulate(p.plan(pts), eps, x0.clone())
    assert float((got - ref).abs().max()) < 5e-6 * float(ref.abs().max())
    sa, sb = p.add_noise_coeffs(pts[0])
    n = torch.randn(4, 8, 8, generator=g)
    assert torch.allclose(o.add_noise(x0.float(), n, torch.tensor([pts[0]])), sa * x0.float() + sb * n, atol=1e-6)


# ------------------------------------------------------------------------------------------------ oracle pipelines (tiny run)
def test_oracle_img2img_and_inpaint_run_and_are_deterministic():
    """Small-latent CPU run of both restated pipelines: shapes, RNG draw order, determinism for a fixed seed."""
    from oracle.pipelines import OraclePipeline, Trace
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from oracle.vae import AutoencoderKL, VAEConfig
    torch.manual_seed(0)
    cfgs = dict(block_out_channels=(32, 64, 64, 64), attention_head_dim=4, cross_attention_dim=32, norm_num_groups=8)
    vae = AutoencoderKL(VAEConfig(block_out_channels=(32, 32, 32, 32), norm_num_groups=8)).eval()
    pe, ne = torch.randn(1, 7, 32), torch.randn(1, 7, 32)
    img = torch.rand(1, 3, 64, 64) * 2 - 1
    for cin, kind in ((4, "pndm"), (9, "ddim")):
        unet = UNet2DConditionModel(UNetConfig(in_channels=cin, **cfgs)).eval()
        pipe = OraclePipeline(unet, vae, kind)
        outs = []
        for _ in range(2):
            tr = Trace()
            g = torch.Generator().manual_seed(42)
            if cin == 4:
                o = pipe.img2img(img, pe, ne, strength=0.5, num_inference_steps=10, guidance_scale=5.0, generator=g, trace=tr)
            else:
                m = torch.zeros(1, 1, 64, 64); m[:, :, 16:32, 16:48] = 1
                o = pipe.inpaint(img, m, pe, ne, strength=0.6, num_inference_steps=10, guidance_scale=5.0, generator=g, trace=tr)
            outs.append(o)
            assert o.shape == (1, 64, 64, 3) and o.dtype == np.uint8
            assert tr.unet_in[0].shape == (2, cin, 8, 8)
        assert (outs[0] == outs[1]).all()
    with pytest.raises(ValueError):
        pipe.inpaint(img, m, pe, ne, strength=1.5)


def test_preprocess_demo_sizes_golden():
    """VaeImageProcessor: (w - w % 8, h - h % 8) LANCZOS; sizes of the reference's data/demo images."""
    gold = json.loads((GOLD / "demo_preprocess.json").read_text())
    sizes = {tuple(v["in_size"]): tuple(v["out_hw"]) for v in gold.values()}
    assert sizes[(500, 333)] == (328, 496) and sizes[(640, 457)] == (456, 640)
    from PIL import Image
    from oracle.pipelines import postprocess_image, preprocess_image, preprocess_mask
    im = Image.fromarray((np.arange(50 * 37 * 3) % 255).astype(np.uint8).reshape(37, 50, 3))
    t = preprocess_image(im)
    assert t.shape == (1, 3, 32, 48) and float(t.min()) >= -1 and float(t.max()) <= 1
    m = preprocess_mask(Image.fromarray(np.full((37, 50), 128, np.uint8)), 512, 512)
    assert m.shape == (1, 1, 512, 512) and set(np.unique(m.numpy())) == {1.0}
    u8 = postprocess_image(torch.tensor([[[[-1.0, 1.0, 0.0, 3.0]]]]).expand(1, 3, 1, 4))
    assert u8[0, 0].tolist() == [[0] * 3, [255] * 3, [128] * 3, [255] * 3]


# ------------------------------------------------------------------------------------------------ metrics
def test_psnr_ssim_known_answers():
    from image_restoration_and_enhancement_b200 import metrics as M
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (32, 40, 3), dtype=np.uint8)
    b = a.copy(); b[0, 0, 0] = (int(b[0, 0, 0]) + 10) % 256
    d = float(int(a[0, 0, 0]) - int(b[0, 0, 0])) ** 2 / a.size
    assert M.psnr(a, b) == pytest.approx(10 * math.log10(255 ** 2 / d), rel=1e-12)
    assert M.psnr(a, a) == float("inf")
    assert M.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    # brute-force SSIM (definition: 7x7 windows fully inside the image, sample covariance) on one channel
    x, y = a[..., 0].astype(np.float64), rng.integers(0, 256, (32, 40)).astype(np.float64)
    C1, C2, vals = (0.01 * 255) ** 2, (0.03 * 255) ** 2, []
    for i in range(3, 32 - 3):
        for j in range(3, 40 - 3):
            wx, wy = x[i - 3:i + 4, j - 3:j + 4], y[i - 3:i + 4, j - 3:j + 4]
            mx, my = wx.mean(), wy.mean()
            vx, vy = wx.var(ddof=1), wy.var(ddof=1)
            cxy = ((wx - mx) * (wy - my)).sum() / 48
            vals.append((2 * mx * my + C1) * (2 * cxy + C2) / ((mx ** 2 + my ** 2 + C1) * (vx + vy + C2)))
    assert M.ssim(x.astype(np.uint8), y.astype(np.uint8), channel_axis=None) == pytest.approx(np.mean(vals), rel=1e-9)
    agg = M.aggregate([1.0, 2.0, 4.0])
    assert agg["mean"] == pytest.approx(7 / 3) and agg["median"] == 2.0 and agg["std"] == pytest.approx(np.std([1, 2, 4]))


# ------------------------------------------------------------------------------------------------ host-side reference API
def test_restoration_pipeline_surface_and_host_glue():
    from PIL import Image
    from image_restoration_and_enhancement_b200.inference import (DEFAULT_PROMPTS, SAMPLING, TASK_MODEL_DIRS,
                                                                  RestorationPipeline)
    import inspect
    sig = inspect.signature(RestorationPipeline.__init__)
    assert list(sig.parameters)[:4] == ["self", "device", "config", "seed"] and sig.parameters["seed"].default == 42
    p = RestorationPipeline(device="cpu", backend="fine_tuned")            # generate_predictions.py:18 passes backend=
    assert set(TASK_MODEL_DIRS) == {"denoise", "sr", "colorize", "inpaint"}
    assert p.prompts == DEFAULT_PROMPTS and p.prompts["denoise"].startswith("clean high quality photo")
    assert SAMPLING["colorize"] == dict(num_inference_steps=30, guidance_scale=7.5, strength=0.75)
    assert SAMPLING["inpaint"] == dict(num_inference_steps=30, guidance_scale=5.0, strength=0.6)
    assert SAMPLING["denoise"]["num_inference_steps"] == 20 and SAMPLING["sr"]["guidance_scale"] == 0
    # mask polarity: < 10 % white => inverted
    m = np.zeros((40, 60), np.uint8); m[:4, :6] = 255
    out = np.array(p._normalize_mask(Image.fromarray(m), (60, 40)))
    assert out[0, 0] == 0 and out[-1, -1] == 255
    m2 = np.zeros((40, 60), np.uint8); m2[:20] = 255
    assert np.array(p._normalize_mask(Image.fromarray(m2), (30, 20))).shape == (20, 30)
    # colour detection and gray -> RGB by channel 0
    col = np.zeros((8, 8, 3), np.uint8); col[..., 0] = 200
    assert p._colorize_prepare(Image.fromarray(col))[1] is True
    gray = np.stack([np.full((8, 8), 7, np.uint8), np.full((8, 8), 9, np.uint8), np.full((8, 8), 9, np.uint8)], 2)
    im, coloured = p._colorize_prepare(Image.fromarray(gray))
    assert not coloured and (np.array(im) == 7).all()
    # > 1 MP inputs are reduced to a 1024 long side
    assert p._sr_limit(Image.new("RGB", (2048, 1024))).size == (1024, 512)
    assert p._sr_limit(Image.new("RGB", (640, 480))).size == (640, 480)
    # auto-mask: nothing to inpaint in a mid-gray image
    assert p._auto_mask_from_image(Image.new("RGB", (64, 64), (120, 120, 120))) is None
    dark = np.full((64, 64, 3), 120, np.uint8); dark[10:40, 10:40] = 0
    assert p._auto_mask_from_image(Image.fromarray(dark)) is not None
    # on a CPU-only host the SD path is unavailable: tasks fall back to classical CV like the reference, and
    # process() keeps the reference's result keys
    res = p.process(Image.new("RGB", (32, 32), (90, 90, 90)), ["denoise", "sr"], sr_scale=2)
    assert set(res) == {"original", "final", "denoised", "super_resolved"} and res["final"].size == (64, 64)
    with pytest.raises(RuntimeError):
        RestorationPipeline(device="cpu", strict=True).denoise(Image.new("RGB", (32, 32)))


def test_pipeline_contract_without_gpu():
    from image_restoration_and_enhancement_b200.pipelines import (StableDiffusionImg2ImgPipeline,
                                                                   StableDiffusionInpaintPipeline)
    with pytest.raises(OSError):
        StableDiffusionImg2ImgPipeline.from_pretrained("/nonexistent/dir", torch_dtype=torch.float16, use_safetensors=True)
    assert StableDiffusionInpaintPipeline._in_channels == 9 and StableDiffusionImg2ImgPipeline._in_channels == 4


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    import ctypes
    from image_restoration_and_enhancement_b200 import _lib, build
    build.build()
    header = (ROOT / "include" / "restoragen.h").read_text()
    declared = set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", header))
    assert declared and declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().rg_version() >= 100
    # the two builds of the same sources: bf16 operands (the product) and the fp16 parity mode
    assert _lib.load().rg_operand_dtype() == _lib.RG_DT_BF16
    lib16 = ctypes.CDLL(str(build.LIB_F16))
    for name in declared:
        assert hasattr(lib16, name), name
    assert lib16.rg_operand_dtype() == _lib.RG_DT_F16
    sizes = {"rg_conv_t": ctypes.sizeof(_lib.RgConv), "rg_attn_t": ctypes.sizeof(_lib.RgAttn)}
    assert sizes["rg_conv_t"] % 8 == 0 and sizes["rg_attn_t"] % 8 == 0


def test_weight_packing_identities():
    from image_restoration_and_enhancement_b200.weights import interleave_geglu, pack_conv, upsample_parity_weights
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    w = torch.randn(6, 4, 3, 3, generator=g)
    assert torch.equal(pack_conv(w)[:, 4:8], w[:, :, 0, 1])
    # upsample + 3x3 conv == four 2x2 parity convs (exact in fp64)
    x = torch.randn(1, 4, 5, 7, generator=g).double(); wd = w.double()
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wd, padding=1)
    out = torch.zeros_like(ref)
    for py, px, wp in upsample_parity_weights(wd):
        k = wp.double().view(6, 2, 2, 4).permute(0, 3, 1, 2)
        xp = F.pad(x, (1 - px, px, 1 - py, py))
        out[:, :, py::2, px::2] = F.conv2d(xp, k)
    assert float((out - ref).abs().max()) < 1e-5
    wg, bg = torch.arange(640.0)[:, None].repeat(1, 2), torch.arange(640.0)
    wi, bi = interleave_geglu(wg, bg)
    # units of 32 rows: 16 value rows followed by their 16 gate rows (GEMM epilogue, csrc/gemm.cu)
    assert bi[:16].tolist() == list(range(16)) and bi[16:32].tolist() == list(range(320, 336))
    assert bi[32:48].tolist() == list(range(16, 32)) and torch.equal(wi[:, 0], bi)


def test_operand_dtype_selection_is_per_process():
    """RESTORAGEN_OPERAND_DTYPE picks the library and the host modules' 16-bit dtype at import (fp16 parity mode)."""
    import os, subprocess, sys
    code = ("from image_restoration_and_enhancement_b200 import _lib, ops; "
            "print(_lib.LIB_PATH.name, ops.OPERAND_DTYPE, _lib.load().rg_operand_dtype())")
    for env, want in (("", "librestoragen.so torch.bfloat16 0"), ("fp16", "librestoragen_f16.so torch.float16 2")):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT),
                           env=dict(os.environ, RESTORAGEN_OPERAND_DTYPE=env))
        assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == want, (r.stdout, r.stderr[-500:])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT),
                       env=dict(os.environ, RESTORAGEN_OPERAND_DTYPE="int8"))
    assert r.returncode != 0 and "RESTORAGEN_OPERAND_DTYPE" in r.stderr
