"""Shared fixtures for the parity tests: seeded rays, config pairs (oracle <-> engine)."""
import os

import numpy as np
import torch

from oracle import mipnerf360 as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def basis_np():
  return np.load(os.path.join(GOLDEN, 'geopoly_basis.npz'))['icosahedron_2'].T.astype(np.float32)  # [3,21]


def make_rays(n, seed=0, near=0.2, far=1e6, scene_radius=1.0, glo=False, n_embed=16):
  """Cameras on a sphere of radius ~1 looking inwards (SURVEY §8d config 2 geometry)."""
  rng = np.random.default_rng(seed)
  cam = rng.normal(size=(n, 3)); cam /= np.linalg.norm(cam, axis=-1, keepdims=True); cam *= scene_radius
  target = rng.normal(size=(n, 3)) * 0.2
  d = target - cam
  d /= np.linalg.norm(d, axis=-1, keepdims=True)
  d *= rng.uniform(0.9, 1.1, size=(n, 1))          # directions are not unit-norm in general
  v = d / np.linalg.norm(d, axis=-1, keepdims=True)
  rays = dict(
      origins=torch.tensor(cam, dtype=torch.float32), directions=torch.tensor(d, dtype=torch.float32),
      viewdirs=torch.tensor(v, dtype=torch.float32),
      radii=torch.tensor(rng.uniform(5e-4, 2e-3, size=(n, 1)), dtype=torch.float32),
      near=torch.full((n, 1), float(near)), far=torch.full((n, 1), float(far)),
      lossmult=torch.ones(n, 1), static_mask=torch.tensor((rng.uniform(size=(n, 1)) < 0.8).astype(np.float32)),
      embed_idx=torch.tensor(rng.integers(0, n_embed, size=(n, 1)), dtype=torch.int32))
  gt = torch.tensor(rng.uniform(size=(n, 3)), dtype=torch.float32)
  return rays, gt


def config_pair(num_levels=2, n_prop=64, n_nerf=128, width=256, nerf_depth=8, prop_depth=4, precision='fp32',
                max_rays=256, glo=0, contract=True, raydist='reciprocal', opaque=True):
  from nerf_hugs_b200.engine import EngineConfig
  warp = 'contract' if contract else None
  ocfg = O.ModelConfig(num_levels=num_levels, num_prop_samples=n_prop, num_nerf_samples=n_nerf,
                       raydist_fn=raydist, opaque_background=opaque, num_glo_features=glo, num_embeddings=16,
                       nerf_mlp=O.MLPConfig(net_depth=nerf_depth, net_width=width, warp_fn=warp),
                       prop_mlp=O.MLPConfig(net_depth=prop_depth, net_width=width, disable_rgb=True, warp_fn=warp))
  ecfg = EngineConfig(num_levels=num_levels, num_prop_samples=n_prop, num_nerf_samples=n_nerf,
                      nerf_depth=nerf_depth, nerf_width=width, prop_depth=prop_depth, prop_width=width,
                      raydist_fn=raydist, nerf_contract=contract, prop_contract=contract,
                      opaque_background=opaque, num_glo_features=glo, num_embeddings=16, precision=precision,
                      max_rays=max_rays)
  return ocfg, ecfg
