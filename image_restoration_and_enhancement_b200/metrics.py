"""Per-image restoration metrics and their bookkeeping (reference ``/root/reference/src/metrics.py``).

``MetricsCalculator.calculate_psnr/ssim`` restate the two scikit-image calls the reference makes
(``src/metrics.py:87`` ``psnr(gt, pred, data_range=255.0)``; ``:95`` ``ssim(gt, pred, data_range=255.0,
channel_axis=2)``) -- scikit-image itself is not installed here -- in float64 numpy/scipy, and ``aggregate``
is the reference's statistics block (``:338-346``).  ``gather_per_image`` is the multi-GPU piece: every rank
contributes its per-image float64 values with their global indices, rank 0 reorders by index and aggregates, so
the result is bit-identical to a single-process run (partial sums are never all-reduced).
LPIPS (``src/metrics.py:67,97-111``): on a CUDA device ``use_lpips=True`` scores with ``lpips.LPIPSB200`` (AlexNet taps on
the tcgen05 conv kernel); the pretrained AlexNet + lin-head files are not available offline, so unless
``lpips_weights=(alexnet_state_dict_path, lpips_alex_pth_path)`` is given the network is the seeded random-init one and
the calculator says so loudly (values unpinned, bookkeeping exact).  On the CPU LPIPS needs ``lpips_fn`` (a callable);
without one a warning is printed, as the reference does when the ``lpips`` package is missing (``:24-29``).

``psnr_ssim_device`` is the GPU metric pass (SURVEY 8f "f2"): the predictions never leave HBM, the kernels in
``csrc/metrics.cu`` reproduce the float64 operation ORDER of scipy's uniform filter and numpy's mean, and the values
returned are bit-identical to ``calculate_psnr`` / ``calculate_ssim`` (tests/test_kernels_gpu.py).
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable

import numpy as np


def load_image(path: Path) -> np.ndarray:
    """RGB uint8 array, decoded the way the reference does (cv2.imread + BGR2RGB, ``src/metrics.py:40-46``)."""
    import cv2
    img = cv2.imread(str(path))
    if img is None:
        raise ValueError(f"Could not load image: {path}")
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def _match_shape(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    if pred.shape != gt.shape:
        import cv2
        pred = cv2.resize(pred, (gt.shape[1], gt.shape[0]))
    return pred


def psnr(gt: np.ndarray, pred: np.ndarray, data_range: float = 255.0) -> float:
    """skimage.metrics.peak_signal_noise_ratio: float64 MSE, 10 log10(R^2 / mse)."""
    a, b = gt.astype(np.float64), pred.astype(np.float64)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10 * np.log10((data_range ** 2) / err))


def _ssim_2d(im1: np.ndarray, im2: np.ndarray, data_range: float, win_size: int = 7, K1: float = 0.01,
             K2: float = 0.03) -> float:
    from scipy.ndimage import uniform_filter
    im1, im2 = im1.astype(np.float64), im2.astype(np.float64)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)                         # sample covariance
    ux, uy = uniform_filter(im1, size=win_size), uniform_filter(im2, size=win_size)
    uxx, uyy, uxy = (uniform_filter(im1 * im1, size=win_size), uniform_filter(im2 * im2, size=win_size),
                     uniform_filter(im1 * im2, size=win_size))
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim(gt: np.ndarray, pred: np.ndarray, data_range: float = 255.0, channel_axis: int | None = 2) -> float:
    """skimage.metrics.structural_similarity (7x7 uniform window, per-channel mean of the cropped SSIM map)."""
    if channel_axis is None or gt.ndim == 2:
        return _ssim_2d(gt, pred, data_range)
    vals = np.empty(gt.shape[channel_axis], dtype=np.float64)
    for c in range(gt.shape[channel_axis]):
        vals[c] = _ssim_2d(np.take(gt, c, axis=channel_axis), np.take(pred, c, axis=channel_axis), data_range)
    return float(vals.mean())


_XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169],
                          [0.019334, 0.119193, 0.950227]])
_D65_WHITE = np.array([0.95047, 1.0, 1.08883])


def rgb2lab(rgb: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab (illuminant D65, observer 2) for float RGB in [0, 1]; keeps the input's float dtype."""
    arr = np.array(rgb, dtype=rgb.dtype if rgb.dtype in (np.float32, np.float64) else np.float64, copy=True)
    m = arr > 0.04045
    arr[m] = np.power((arr[m] + 0.055) / 1.055, 2.4)
    arr[~m] /= 12.92
    xyz = arr @ _XYZ_FROM_RGB.T.astype(arr.dtype)
    xyz = xyz / _D65_WHITE.astype(arr.dtype)
    m = xyz > 0.008856
    xyz[m] = np.cbrt(xyz[m])
    xyz[~m] = 7.787 * xyz[~m] + 16.0 / 116.0
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    return np.stack([116.0 * y - 16.0, 500.0 * (x - y), 200.0 * (y - z)], axis=-1)


def psnr_ssim_device(pred, gt, data_range: float = 255.0, K1: float = 0.01, K2: float = 0.03):
    """PSNR and SSIM of u8 image batches resident on the GPU: ``pred``, ``gt`` torch.uint8 CUDA tensors [N,H,W,C].
    Returns two lists of Python floats, bit-identical to ``psnr(gt[i], pred[i])`` / ``ssim(gt[i], pred[i])``.
    The scalar tails (mse -> dB, chunk sums -> mean) are the same numpy float64 expressions the CPU functions use."""
    import ctypes as C
    import torch
    from . import _lib
    if not (pred.is_cuda and gt.is_cuda) or pred.dtype != torch.uint8 or gt.dtype != torch.uint8:
        raise ValueError("psnr_ssim_device needs torch.uint8 CUDA tensors")
    if pred.shape != gt.shape or pred.dim() != 4:
        raise ValueError(f"shape mismatch or not [N,H,W,C]: {tuple(pred.shape)} vs {tuple(gt.shape)}")
    lib = _lib.load()
    pred, gt = pred.contiguous(), gt.contiguous()
    N, H, W, Cc = pred.shape
    chunks = lib.rg_metrics_ssim_chunks(H, W)
    if chunks < 1:
        raise ValueError(f"SSIM on the GPU needs 7 <= H, 7 <= W <= 8198 (got {H}x{W})")
    dev = pred.device
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    sse = torch.empty((N,), dtype=torch.int64, device=dev)
    smap = torch.empty((N, Cc, H - 6, W - 6), dtype=torch.float64, device=dev)
    csum = torch.empty((N, Cc, chunks), dtype=torch.float64, device=dev)
    NP = 7 ** 2
    cov_norm = NP / (NP - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    _lib.check(lib.rg_metrics_sse_u8(pred.data_ptr(), gt.data_ptr(), N, H * W * Cc, sse.data_ptr(), stream), "rg_metrics_sse_u8")
    _lib.check(lib.rg_metrics_ssim_u8(pred.data_ptr(), gt.data_ptr(), N, H, W, Cc, C1, C2, cov_norm, smap.data_ptr(),
                                      csum.data_ptr(), stream), "rg_metrics_ssim_u8")
    sse_h = sse.cpu().numpy()
    csum_h = csum.cpu().numpy()
    count = (H - 6) * (W - 6)
    psnrs, ssims = [], []
    for n in range(N):
        err = np.float64(int(sse_h[n])) / (H * W * Cc)              # == np.mean of the exact integer squares
        with np.errstate(divide="ignore"):
            psnrs.append(float(10 * np.log10((data_range ** 2) / err)))
        vals = np.empty(Cc, dtype=np.float64)
        for c in range(Cc):
            t = 0.0
            for v in csum_h[n, c]:
                t += float(v)                                        # numpy adds its buffer chunks in order
            vals[c] = np.float64(t) / count
        ssims.append(float(vals.mean()))
    return psnrs, ssims


class MetricsCalculator:
    """``device`` "cuda" (any CUDA device string) scores PSNR / SSIM with the kernels of csrc/metrics.cu -- same float64
    bits as "cpu" -- and raises if librestoragen.so or a GPU is missing (no silent fallback)."""

    def __init__(self, use_lpips: bool = True, use_fid: bool = False, device: str = "cpu",
                 lpips_fn: Callable[[np.ndarray, np.ndarray], float] | None = None, lpips_weights=None, lpips_seed: int = 0):
        import sys
        self.lpips_fn = lpips_fn
        self.use_fid = False              # FID needs Inception weights that are not available offline
        self.device = device
        self._gpu = str(device).startswith("cuda")
        self._lpips_model = None
        if use_lpips and lpips_fn is None and self._gpu:
            from . import lpips as _lp
            if lpips_weights is not None:
                import torch
                a, l = (torch.load(str(f), map_location="cpu") for f in lpips_weights)
                sd = _lp.load_lpips_state_dict(a, l)
            else:
                print("Warning: pretrained LPIPS (AlexNet) weights are not available offline -- using the seeded random-init "
                      f"network (seed {lpips_seed}); LPIPS VALUES are not comparable with published ones", file=sys.stderr)
                sd = _lp.random_lpips_state_dict(lpips_seed)
            self._lpips_model = _lp.LPIPSB200(sd, device=str(device))
        elif use_lpips and lpips_fn is None:
            print("Warning: LPIPS not available on the CPU path (no lpips_fn given); use device='cuda'", file=sys.stderr)
        self.use_lpips = use_lpips and (lpips_fn is not None or self._lpips_model is not None)

    def _device_pair(self, pred: np.ndarray, gt: np.ndarray) -> tuple[float, float]:
        import torch
        pred = _match_shape(pred, gt)
        p3, g3 = (pred[..., None], gt[..., None]) if gt.ndim == 2 else (pred, gt)
        p, s = psnr_ssim_device(torch.from_numpy(np.ascontiguousarray(p3)[None]).to(self.device),
                                torch.from_numpy(np.ascontiguousarray(g3)[None]).to(self.device))
        return p[0], s[0]

    def calculate_psnr(self, pred: np.ndarray, gt: np.ndarray) -> float:
        return psnr(gt, _match_shape(pred, gt), data_range=255.0)

    def calculate_ssim(self, pred: np.ndarray, gt: np.ndarray) -> float:
        return ssim(gt, _match_shape(pred, gt), data_range=255.0, channel_axis=2 if gt.ndim == 3 else None)

    def calculate_lpips(self, pred: np.ndarray, gt: np.ndarray):
        if not self.use_lpips:
            return None
        pred = _match_shape(pred, gt)
        if self.lpips_fn is not None:
            return float(self.lpips_fn(pred, gt))
        import torch
        to3 = lambda a: np.ascontiguousarray(np.repeat(a[..., None], 3, axis=2) if a.ndim == 2 else a)
        return self._lpips_model(torch.from_numpy(to3(pred)[None]).to(self.device), torch.from_numpy(to3(gt)[None]).to(self.device))[0]

    def calculate_delta_e(self, pred: np.ndarray, gt: np.ndarray, use_delta_e2000: bool = False) -> float:
        """Mean Delta-E 76 in CIELAB (reference ``src/metrics.py:113-148``; its ``use_delta_e2000`` branch computes the
        same Euclidean distance).  ``rgb2lab`` restates skimage.color (sRGB companding, D65 / 2-degree white point) in
        the float32 arithmetic the reference feeds it."""
        pred = _match_shape(pred, gt)
        d = rgb2lab(pred.astype(np.float32) / 255.0) - rgb2lab(gt.astype(np.float32) / 255.0)
        return float(np.mean(np.sqrt(np.sum(d ** 2, axis=2))))

    def calculate_all(self, pred: np.ndarray, gt: np.ndarray) -> dict:
        if self._gpu:
            p, s = self._device_pair(pred, gt)
            out = {"psnr": p, "ssim": s}
        else:
            out = {"psnr": self.calculate_psnr(pred, gt), "ssim": self.calculate_ssim(pred, gt)}
        if self.use_lpips:
            out["lpips"] = self.calculate_lpips(pred, gt)
        return out


def aggregate(values: list[float]) -> dict:
    """mean / std (ddof 0) / min / max / median in float64, as ``src/metrics.py:338-346``."""
    return {"mean": np.mean(values), "std": np.std(values), "min": np.min(values), "max": np.max(values),
            "median": np.median(values)}


def summarize(task: str, per_image: dict[str, list[float]], num_samples: int) -> dict:
    return {"task": task, "num_samples": num_samples,
            "metrics": {k: aggregate(v) for k, v in per_image.items() if v}}


def gather_per_image(indices: list[int], metrics: dict[str, list[float]], group=None) -> dict[str, list[float]] | None:
    """All ranks pass the global indices of the images they evaluated and the per-image values; rank 0 gets the
    values of every image in global-index order (others get None).  Works on NCCL (device tensors) and gloo."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = np.argsort(np.asarray(indices, dtype=np.int64), kind="stable")
        return {k: [v[i] for i in order] for k, v in metrics.items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    keys = sorted(metrics)
    n_local = torch.tensor([len(indices)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(int(c) for c in counts))
    # one float64 row per image: [global index, metric values...]; exact for indices < 2^53
    rows = torch.full((n_max, 1 + len(keys)), float("nan"), dtype=torch.float64, device=dev)
    if indices:
        local = np.column_stack([np.asarray(indices, dtype=np.float64)] + [np.asarray(metrics[k], dtype=np.float64) for k in keys])
        rows[:len(indices)] = torch.from_numpy(local).to(dev)
    gathered = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows, group=group)
    if rank != 0:
        return None
    allrows = np.concatenate([g[:int(c)].cpu().numpy() for g, c in zip(gathered, counts)], axis=0)
    order = np.argsort(allrows[:, 0].astype(np.int64), kind="stable")
    allrows = allrows[order]
    return {k: [float(x) for x in allrows[:, 1 + j]] for j, k in enumerate(keys)}


def evaluate_task(pred_dir: Path, gt_dir: Path, task_name: str = "denoise", use_lpips: bool = True,
                  use_fid: bool = False, device: str = "cpu") -> dict:
    """Directory-based evaluation with the reference's file matching (``src/metrics.py:238-348``)."""
    calc = MetricsCalculator(use_lpips=use_lpips, use_fid=use_fid, device=device)
    exts = {".jpg", ".jpeg", ".png"}
    pred_files = sorted(f for f in Path(pred_dir).iterdir() if f.suffix.lower() in exts)
    pairs = []
    for pf in pred_files:
        gf = Path(gt_dir) / pf.name
        if not gf.exists():
            for ext in (".jpg", ".jpeg", ".png"):
                alt = Path(gt_dir) / (pf.stem + ext)
                if alt.exists():
                    gf = alt
                    break
        if gf.exists():
            pairs.append((pf, gf))
    if not pairs:
        raise ValueError(f"No matching files found between {pred_dir} and {gt_dir}")
    per_image: dict[str, list[float]] = {"psnr": [], "ssim": []}
    for pf, gf in pairs:
        try:
            m = calc.calculate_all(load_image(pf), load_image(gf))
        except Exception as e:             # the reference skips unreadable pairs
            print(f"Error processing {pf.name}: {e}")
            continue
        for k, v in m.items():
            if v is not None:
                per_image.setdefault(k, []).append(v)
    return summarize(task_name, per_image, len(pairs))


def print_results(results: dict) -> None:
    """Pretty print evaluation results (reference ``src/metrics.py:351-365``)."""
    print(f"\n{'=' * 60}")
    print(f"Evaluation Results: {results['task']}")
    print(f"{'=' * 60}")
    print(f"Number of samples: {results['num_samples']}")
    print("\nMetrics:")
    for metric_name, stats in results["metrics"].items():
        print(f"\n  {metric_name.upper()}:")
        print(f"    Mean:   {stats['mean']:.4f} \u00b1 {stats['std']:.4f}")
        print(f"    Median: {stats['median']:.4f}")
        print(f"    Range:  [{stats['min']:.4f}, {stats['max']:.4f}]")
    print(f"\n{'=' * 60}\n")
