"""nerf_hugs_b200 — B200-native per-ray volume-rendering path of cnhaox/NeRF-HuGS.

`nerf_hugs_b200.engine` is the ctypes host over libhugs_b200.so (include/hugs_b200.h);
`nerf_hugs_b200.internal` re-creates the call surface of the reference's MipNeRF360/internal
(configs, utils, models, train_utils) on top of it.  Importing the engine without the built
shared library raises: there is no CPU / PyTorch fallback for the compute path.
"""
__version__ = '0.1.0'
