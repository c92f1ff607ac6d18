"""The hand-off between "predict" and "evaluate" (SURVEY 8f "f1").

The reference runs two scripts that talk through image files named after the inputs:

* ``scripts/generate_predictions.py:60-84`` -- ``Image.open(p).convert("RGB")`` (PIL decode) -> pipeline ->
  ``result["final"].save(output_dir / p.name)``: PIL encodes with its defaults, i.e. baseline JPEG quality 75 with 4:2:0
  chroma subsampling for ``.jpg`` inputs (denoise / sr_x4 / inpaint) and lossless PNG for ``.png`` inputs (colorize,
  whose grayscale inputs are written as ``<stem>.png``, ``scripts/make_synthetic_pairs.py:181-182``).
* ``scripts/evaluate_model.py`` -> ``src/metrics.py:40-46`` reloads prediction and ground truth with ``cv2.imread`` +
  BGR->RGB, ``cv2.resize`` of the prediction on a shape mismatch (``:85-86``).

So the u8 array that gets scored is NOT the pipeline's output: it is ``cv2.decode(PIL.encode(output))``, and the ground
truth is ``cv2.decode`` of whatever ``cv2.imwrite`` (JPEG quality 95) stored.  "Bit-exact metric bookkeeping" is only
defined once that codec round trip is part of the path.  This module provides it twice:

* on disk, with the reference's file layout (``save_prediction`` / ``load_image`` / dataset writers), and
* in memory (``roundtrip_prediction`` / ``roundtrip_dataset_image``): the same encoders and decoders applied to byte
  buffers, no file system -- tests/test_handoff_cpu.py proves the arrays are identical to the on-disk ones, and that PNG
  hand-offs are the identity so they are skipped altogether.

Nothing here touches the GPU; the codecs are PIL and OpenCV, exactly the libraries the reference uses for them.
"""
from __future__ import annotations

import io
from pathlib import Path

import numpy as np
from PIL import Image

IMG_EXTS = (".jpg", ".jpeg", ".png")                     # scripts/make_synthetic_pairs.py:17, src/metrics.py:306
# directory name under pairs/ and predictions/ -> RestorationPipeline task (scripts/generate_predictions.py:21-38)
TASK_DIRS = {"denoise": "denoise", "sr_x4": "sr", "colorize": "colorize", "inpaint": "inpaint"}


def _suffix(name) -> str:
    suf = Path(str(name)).suffix.lower()
    if suf not in IMG_EXTS:
        raise ValueError(f"unsupported image extension: {name}")
    return suf


def input_name(task_dir: str, index: int) -> str:
    """File name of work item ``index`` in ``pairs/<task_dir>/<split>/input``: colorize inputs are PNG, the rest keep the
    source image's ``.jpg`` (``scripts/make_synthetic_pairs.py:171-195``)."""
    return f"{index:06d}.png" if task_dir == "colorize" else f"{index:06d}.jpg"


def gt_name(task_dir: str, index: int) -> str:
    return f"{index:06d}.jpg"                            # ground truth always keeps the source name (:172,177,183,195)


# ---------------------------------------------------------------------------------------------------------------------
# prediction side: PIL encodes, cv2 decodes
def encode_prediction(img: Image.Image, name) -> bytes:
    """The bytes ``img.save(path)`` writes for a path with this name (format chosen from the extension, PIL defaults)."""
    fmt = Image.registered_extensions()[_suffix(name)]
    buf = io.BytesIO()
    img.save(buf, format=fmt)
    return buf.getvalue()


def decode_cv2(data: bytes) -> np.ndarray:
    """``src/metrics.py:load_image`` on a byte buffer: cv2 decode (IMREAD_COLOR) + BGR->RGB."""
    import cv2
    arr = cv2.imdecode(np.frombuffer(data, dtype=np.uint8), cv2.IMREAD_COLOR)
    if arr is None:
        raise ValueError("could not decode image buffer")
    return cv2.cvtColor(arr, cv2.COLOR_BGR2RGB)


def decode_pil(data: bytes, mode: str = "RGB") -> Image.Image:
    """``Image.open(path).convert(mode)`` (``scripts/generate_predictions.py:68,74``) on a byte buffer."""
    return Image.open(io.BytesIO(data)).convert(mode)


def is_lossless(name) -> bool:
    return _suffix(name) == ".png"


def roundtrip_prediction(img: Image.Image, name) -> np.ndarray:
    """The RGB u8 array the evaluator would score for this prediction, without touching the file system.
    PNG hand-offs are the identity on RGB u8 images and are skipped."""
    if is_lossless(name) and img.mode == "RGB":
        return np.array(img)
    return decode_cv2(encode_prediction(img, name))


def save_prediction(img: Image.Image, path) -> None:
    """``result["final"].save(output_path)`` (``scripts/generate_predictions.py:83-84``)."""
    _suffix(path)
    img.save(str(path))


def load_image(path) -> np.ndarray:
    from .metrics import load_image as _load
    return _load(Path(path))


# ---------------------------------------------------------------------------------------------------------------------
# dataset side: cv2 encodes (scripts/make_synthetic_pairs.py:171-195 cv2.imwrite, defaults: JPEG quality 95)
def encode_dataset_image(rgb_or_gray: np.ndarray, name) -> bytes:
    import cv2
    arr = rgb_or_gray if rgb_or_gray.ndim == 2 else cv2.cvtColor(rgb_or_gray, cv2.COLOR_RGB2BGR)
    ok, enc = cv2.imencode(_suffix(name), arr)
    if not ok:
        raise ValueError(f"cv2.imencode failed for {name}")
    return enc.tobytes()


def roundtrip_dataset_image(arr: np.ndarray, name, reader: str) -> np.ndarray | Image.Image:
    """What a consumer sees of a dataset image written by ``cv2.imwrite``: ``reader`` "pil" (the predictor's
    ``Image.open(...).convert("RGB")``, returns a PIL image), "pil_l" (the mask, ``.convert("L")``) or "cv2" (the
    evaluator's ``load_image``, returns an RGB array)."""
    data = encode_dataset_image(arr, name)
    if reader == "cv2":
        return decode_cv2(data)
    if reader == "pil":
        return decode_pil(data, "RGB")
    if reader == "pil_l":
        return decode_pil(data, "L")
    raise ValueError(f"unknown reader {reader}")


def write_pairs(root, task_dir: str, index: int, pair: dict, split: str = "test") -> None:
    """One work item in the reference's ``pairs/<task>/<split>/{input,gt[,mask]}`` layout, written with ``cv2.imwrite``
    like ``scripts/make_synthetic_pairs.py:132-140,171-195``.  ``pair`` is ``synth.make_pair``'s dict (RGB arrays)."""
    import cv2
    base = Path(root) / task_dir / split
    for sub in ("input", "gt") + (("mask",) if "mask" in pair else ()):
        (base / sub).mkdir(parents=True, exist_ok=True)
    inp = pair["input"]
    if task_dir == "colorize":
        inp = inp[:, :, 0]                                # single-channel L image, as :181-182 writes it
    cv2.imwrite(str(base / "input" / input_name(task_dir, index)), inp if inp.ndim == 2 else cv2.cvtColor(inp, cv2.COLOR_RGB2BGR))
    cv2.imwrite(str(base / "gt" / gt_name(task_dir, index)), cv2.cvtColor(pair["gt"], cv2.COLOR_RGB2BGR))
    if "mask" in pair:
        cv2.imwrite(str(base / "mask" / input_name(task_dir, index)), pair["mask"])
