"""restoragen-b200: B200-native (sm_100a) Stable-Diffusion sampling loop behind RestoraGen's RestorationPipeline.

The compute path is librestoragen.so (hand-written CUDA, C ABI in include/restoragen.h); this package is the
Python host side that mirrors the reference's interfaces (src/inference.py, src/metrics.py).
"""
__version__ = "0.1.0"
