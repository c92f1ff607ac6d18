"""The predict -> evaluate hand-off (SURVEY 8f "f1"): the in-memory codec round trip yields exactly the arrays the
reference's two-script, file-based flow scores (scripts/generate_predictions.py:83-84 PIL save by input name ->
src/metrics.py:40-46 cv2.imread), and the sharded sweep gives identical metrics either way."""
import numpy as np
import pytest
from PIL import Image

from image_restoration_and_enhancement_b200 import handoff, metrics, sweep, synth


def _img(i=0, size=96):
    return synth.make_pair("denoise", i, size, size)


@pytest.mark.parametrize("name", ["000001.jpg", "000001.jpeg", "000001.png"])
def test_prediction_roundtrip_memory_equals_disk(tmp_path, name):
    pred = Image.fromarray(_img(1)["input"])
    handoff.save_prediction(pred, tmp_path / name)
    on_disk = handoff.load_image(tmp_path / name)
    in_mem = handoff.roundtrip_prediction(pred, name)
    assert on_disk.dtype == np.uint8 and np.array_equal(on_disk, in_mem)
    assert (tmp_path / name).read_bytes() == handoff.encode_prediction(pred, name)


def test_png_handoff_is_identity_and_jpeg_is_not():
    arr = _img(2)["input"]
    assert np.array_equal(handoff.roundtrip_prediction(Image.fromarray(arr), "x.png"), arr)
    assert np.array_equal(handoff.decode_cv2(handoff.encode_prediction(Image.fromarray(arr), "x.png")), arr)
    jpg = handoff.roundtrip_prediction(Image.fromarray(arr), "x.jpg")
    assert jpg.shape == arr.shape and not np.array_equal(jpg, arr)
    gt = _img(2)["gt"]
    # the codec moves the metric: scoring the raw array is not what the reference's evaluator does for .jpg tasks
    assert metrics.psnr(gt, jpg) != metrics.psnr(gt, arr)


@pytest.mark.parametrize("task", ["denoise", "sr", "colorize", "inpaint"])
def test_dataset_roundtrip_memory_equals_disk(tmp_path, task):
    tdir = sweep.TASK_DIR[task]
    pair = synth.make_pair(task, 3, 64, 64)
    handoff.write_pairs(tmp_path, tdir, 3, pair)
    base = tmp_path / tdir / "test"
    nm = handoff.input_name(tdir, 3)
    assert nm.endswith(".png") == (task == "colorize")
    disk_in = np.array(Image.open(base / "input" / nm).convert("RGB"))
    src = pair["input"][:, :, 0] if task == "colorize" else pair["input"]
    assert np.array_equal(disk_in, np.array(handoff.roundtrip_dataset_image(src, nm, "pil")))
    if task == "colorize":                                  # PNG: the predictor sees the gray image exactly
        assert np.array_equal(disk_in, pair["input"])
    disk_gt = handoff.load_image(base / "gt" / handoff.gt_name(tdir, 3))
    assert np.array_equal(disk_gt, handoff.roundtrip_dataset_image(pair["gt"], handoff.gt_name(tdir, 3), "cv2"))
    if task == "inpaint":
        disk_mask = np.array(Image.open(base / "mask" / nm).convert("L"))
        assert np.array_equal(disk_mask, np.array(handoff.roundtrip_dataset_image(pair["mask"], nm, "pil_l")))


def test_unsupported_extension_is_an_error():
    with pytest.raises(ValueError):
        handoff.roundtrip_prediction(Image.fromarray(_img(0)["input"]), "x.bmp")
    with pytest.raises(ValueError):
        handoff.decode_cv2(b"not an image")


class _BlurPipe:
    """Stand-in for RestorationPipeline on CPU: a deterministic "restoration" (3x3 box blur)."""

    def process_batch(self, images, task, masks=None, **kw):
        import cv2
        if task == "inpaint":
            assert masks is not None and all(m.mode == "L" for m in masks)
        return [Image.fromarray(cv2.blur(np.array(im.convert("RGB")), (3, 3))) for im in images]


@pytest.mark.parametrize("task", ["denoise", "colorize", "inpaint"])
def test_sweep_disk_flow_equals_memory_flow(tmp_path, task):
    pipe = _BlurPipe()
    idx_d, disk, _ = sweep.run_task(pipe, task, 5, size=64, batch=2, handoff_mode="disk", metrics_backend="cpu", workdir=tmp_path)
    idx_m, mem, _ = sweep.run_task(pipe, task, 5, size=64, batch=2, handoff_mode="memory", metrics_backend="cpu")
    assert idx_d == idx_m == [0, 1, 2, 3, 4]
    assert disk == mem                                      # float64 equality, value by value
    _, serial, _ = sweep.run_task(pipe, task, 5, size=64, batch=2, handoff_mode="memory", metrics_backend="cpu", overlap=False)
    assert serial == mem                                    # the host/GPU software pipeline does not change a value
    # ... and equals the reference's directory-based evaluation of the files just written
    ev = metrics.evaluate_task(tmp_path / "predictions" / sweep.TASK_DIR[task] / "test",
                               tmp_path / "pairs" / sweep.TASK_DIR[task] / "test" / "gt", task, use_lpips=False)
    assert ev["num_samples"] == 5
    assert ev["metrics"]["psnr"]["mean"] == np.mean(mem["psnr"]) and ev["metrics"]["ssim"]["median"] == np.median(mem["ssim"])
    _, raw, _ = sweep.run_task(pipe, task, 5, size=64, batch=2, handoff_mode="none", metrics_backend="cpu")
    assert raw != mem                                       # codecs on both sides change what is scored


class _StubSDPipeline:
    """Quacks like the object scripts/train_denoising.py hands to run_validation: .to, .unet/.vae/.text_encoder with
    eval/train, settable safety attributes, __call__ -> .images."""

    class _Mod:
        def __init__(self):
            self.mode = "train"

        def eval(self):
            self.mode = "eval"
            return self

        def train(self, mode=True):
            self.mode = "train"
            return self

    def __init__(self):
        self.unet, self.vae, self.text_encoder = self._Mod(), self._Mod(), self._Mod()
        self.safety_checker, self.feature_extractor, self.requires_safety_checker = object(), object(), True
        self.calls = []

    def to(self, device):
        return self

    def __call__(self, prompt, image, strength, num_inference_steps, guidance_scale, **kw):
        import cv2
        self.calls.append((prompt, strength, num_inference_steps, guidance_scale, self.unet.mode))
        out = cv2.blur(np.array(image), (3, 3))
        return type("Out", (), {"images": [Image.fromarray(out)]})()


def test_run_validation_bookkeeping_cpu(tmp_path):
    """validation.run_validation (mirror of scripts/train_denoising.py:328-520) with a stand-in pipeline: sample
    selection, call parameters, Y-channel metrics, sigma buckets, comparison strips, module modes."""
    import cv2
    import torch
    from image_restoration_and_enhancement_b200 import validation

    def item(i):
        gt = torch.from_numpy(synth.clean_image(i, 64, 64)).permute(2, 0, 1).float() / 127.5 - 1.0
        return {"input": (gt + 0.04 * torch.randn(gt.shape, generator=torch.Generator().manual_seed(i))).clamp(-1, 1),
                "gt": gt, "sigma": 5.2 + i}
    ds = [item(i) for i in range(7)]
    pipe, trained = _StubSDPipeline(), _StubSDPipeline._Mod()
    out = validation.run_validation(2, ds, pipe, trained, tmp_path, num_samples=4, device="cpu")
    assert pipe.unet is trained and trained.mode == "train"                # swapped in, evaluated in eval mode, flipped back
    assert [c[4] for c in pipe.calls] == ["eval"] * 4
    assert all(c[:4] == (validation.VALIDATION_PROMPT, 0.3, 20, 5.0) for c in pipe.calls)
    assert pipe.safety_checker is None and pipe.requires_safety_checker is False
    assert out["num_samples"] == 4 and sorted(out["by_sigma"]) == [5, 7, 9, 11]     # indices 0, 2, 4, 6
    files = sorted(p.name for p in (tmp_path / "val_samples").glob("*.png"))
    assert files == [f"epoch_3_sample_{k + 1}_idx{i}.png" for k, i in enumerate((0, 2, 4, 6))]
    assert Image.open(tmp_path / "val_samples" / files[0]).size == (3 * 64, 64)      # input | result | gt
    # first sample recomputed by hand
    inp, gt = validation._to_u8(ds[0]["input"]), validation._to_u8(ds[0]["gt"])
    res = cv2.blur(inp, (3, 3))
    assert out["per_image"]["psnr"][0] == metrics.psnr(gt, res) and out["per_image"]["ssim"][0] == metrics.ssim(gt, res)
    y = lambda a: cv2.cvtColor(a, cv2.COLOR_RGB2YCrCb)[:, :, 0]
    assert out["by_sigma"][5]["psnr_y"] == metrics.psnr(y(gt), y(res))
    assert validation.run_validation(0, [], pipe, None, tmp_path) is None


def test_gray_prediction_png_comes_back_as_rgb():
    """cv2.imread(IMREAD_COLOR) replicates a single-channel PNG to three channels; the memory path must do the same."""
    g = _img(4)["input"][:, :, 0]
    out = handoff.roundtrip_prediction(Image.fromarray(g, mode="L"), "x.png")
    assert out.shape == g.shape + (3,) and all(np.array_equal(out[:, :, c], g) for c in range(3))


def test_gpu_metric_backend_has_no_silent_fallback():
    """Asking for the GPU metric pass on a machine without CUDA (this test runs on CPU) must raise, not quietly use numpy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    calc = metrics.MetricsCalculator(use_lpips=False, device="cuda")
    a = _img(5)
    with pytest.raises(Exception):
        calc.calculate_all(a["input"], a["gt"])
    with pytest.raises(ValueError):
        metrics.psnr_ssim_device(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), torch.zeros((1, 8, 8, 3), dtype=torch.uint8))
