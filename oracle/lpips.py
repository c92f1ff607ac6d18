"""fp32 PyTorch restatement of ``lpips.LPIPS(net='alex')`` (lpips 0.1.x, the call the reference makes at
``/root/reference/src/metrics.py:67`` and ``:97-111``).  Test infrastructure only (see ``oracle/__init__.py``).

The ``lpips`` package is a ``requirements.txt:13`` dependency (``lpips>=0.1``) that is neither vendored nor installed here;
its published graph is restated: ``ScalingLayer`` -> torchvision AlexNet ``features`` split at the five ReLUs ->
``normalize_tensor`` (``x / (sqrt(sum_c x^2) + 1e-10)``) -> squared difference -> ``NetLinLayer`` (1x1 conv, no bias) ->
``spatial_average`` -> sum over levels.  PARITY UNPINNED: no pretrained weights and no reference outputs exist offline;
state-dict keys follow torchvision (``features.N``) and lpips (``linK.model.1.weight``) so the real files would load.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

SHIFT = (-0.030, -0.088, -0.188)
SCALE = (0.458, 0.448, 0.450)


class LPIPSAlex(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.ModuleDict({
            "0": nn.Conv2d(3, 64, 11, stride=4, padding=2), "3": nn.Conv2d(64, 192, 5, padding=2),
            "6": nn.Conv2d(192, 384, 3, padding=1), "8": nn.Conv2d(384, 256, 3, padding=1),
            "10": nn.Conv2d(256, 256, 3, padding=1)})
        self.lins = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in (64, 192, 384, 256, 256)])
        self.register_buffer("shift", torch.tensor(SHIFT)[None, :, None, None])
        self.register_buffer("scale", torch.tensor(SCALE)[None, :, None, None])

    def load_lpips_state_dict(self, sd: dict):
        own = {}
        for k, v in sd.items():
            if k.startswith("features."):
                own[k] = v
            elif k.startswith("lin"):
                own[f"lins.{k[3]}.weight"] = v
        missing, unexpected = self.load_state_dict(own, strict=False)
        assert not unexpected and set(missing) <= {"shift", "scale"}, (missing, unexpected)
        return self

    def taps(self, x):
        f = self.features
        r1 = F.relu(f["0"](x))
        r2 = F.relu(f["3"](F.max_pool2d(r1, 3, 2)))
        r3 = F.relu(f["6"](F.max_pool2d(r2, 3, 2)))
        r4 = F.relu(f["8"](r3))
        r5 = F.relu(f["10"](r4))
        return [r1, r2, r3, r4, r5]

    @staticmethod
    def _normalize(x, eps=1e-10):
        return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)

    @torch.no_grad()
    def forward(self, in0: torch.Tensor, in1: torch.Tensor) -> torch.Tensor:
        """in0, in1: f32 [N,3,H,W] in [-1, 1] (``preprocess_for_lpips``) -> [N,1,1,1]."""
        t0, t1 = self.taps((in0 - self.shift) / self.scale), self.taps((in1 - self.shift) / self.scale)
        val = 0
        for a, b, lin in zip(t0, t1, self.lins):
            d = (self._normalize(a) - self._normalize(b)) ** 2
            val = val + lin(d).mean(dim=(2, 3), keepdim=True)
        return val


def preprocess_for_lpips(img_u8):
    """``src/metrics.py:48-54``: uint8 HWC -> f32 [1,3,H,W] in [-1, 1]."""
    import numpy as np
    t = torch.from_numpy(img_u8.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    return t * 2.0 - 1.0
