"""``diffusers`` shim: lets the UNMODIFIED reference ``src/inference.py`` (and any other caller that does
``from diffusers import StableDiffusionImg2ImgPipeline, StableDiffusionInpaintPipeline``, reference
``src/inference.py:38-42``) run on the B200 kernels.

Put this directory's parent (``<repo>/shims``) on ``sys.path`` / ``PYTHONPATH`` ahead of site-packages:

    PYTHONPATH=/path/to/repo/shims:/path/to/repo python app.py

Only the two pipeline classes the reference imports are provided; they keep diffusers' construction, placement,
attribute, call and return contract (see image_restoration_and_enhancement_b200/pipelines.py).  Scheduler classes are
exported as well because reference checkpoints name them in ``model_index.json``.
"""
from image_restoration_and_enhancement_b200.pipelines import (  # noqa: F401
    StableDiffusionImg2ImgPipeline,
    StableDiffusionInpaintPipeline,
)
from image_restoration_and_enhancement_b200.schedulers import DDIMScheduler, PNDMScheduler  # noqa: F401

__version__ = "0.35.2+restoragen.b200"
__all__ = ["StableDiffusionImg2ImgPipeline", "StableDiffusionInpaintPipeline", "PNDMScheduler", "DDIMScheduler"]
