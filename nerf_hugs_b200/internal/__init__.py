"""Call surface of the reference's MipNeRF360/internal package, re-hosted over libhugs_b200.so.

Module and symbol names follow /root/reference/MipNeRF360/internal so that train.py / eval.py /
render.py keep working with `from nerf_hugs_b200.internal import configs, models, train_utils, utils`.
"""
