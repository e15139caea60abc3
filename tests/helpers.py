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
                max_rays=256, glo=0, contract=True, raydist='reciprocal', opaque=True, max_deg=12, ray_shape='cone',
                nerf_width=None):
  from nerf_hugs_b200.engine import EngineConfig
  warp = 'contract' if contract else None
  ocfg = O.ModelConfig(num_levels=num_levels, num_prop_samples=n_prop, num_nerf_samples=n_nerf,
                       raydist_fn=raydist, opaque_background=opaque, num_glo_features=glo, num_embeddings=16,
                       ray_shape=ray_shape,
                       nerf_mlp=O.MLPConfig(net_depth=nerf_depth, net_width=nerf_width or width, warp_fn=warp, max_deg_point=max_deg),
                       prop_mlp=O.MLPConfig(net_depth=prop_depth, net_width=width, disable_rgb=True, warp_fn=warp,
                                            max_deg_point=max_deg))
  ecfg = EngineConfig(num_levels=num_levels, num_prop_samples=n_prop, num_nerf_samples=n_nerf,
                      nerf_depth=nerf_depth, nerf_width=nerf_width or width, prop_depth=prop_depth, prop_width=width,
                      raydist_fn=raydist, nerf_contract=contract, prop_contract=contract,
                      opaque_background=opaque, num_glo_features=glo, num_embeddings=16, precision=precision,
                      max_rays=max_rays, max_deg_point=max_deg, ray_shape=ray_shape)
  return ocfg, ecfg


# ----------------------------------------------------------------------------------------------
# Model-level golden cases: the reference's own Model.__call__ was run on these inputs by
# tests/golden/make_golden_mipnerf360.py (fixtures: tests/golden/mip360_model.npz).
# ----------------------------------------------------------------------------------------------
GOLDEN_MODEL_CASES = {
    # SURVEY §8d config A geometry (360.gin: contract + reciprocal spacing, opaque background), deterministic sampling
    'A': dict(n=24, seed=21, near=0.2, far=1e6, levels=2, n_prop=64, n_nerf=128, width=256, nerf_depth=8, prop_depth=4,
              glo=0, contract=True, raydist='reciprocal', opaque=True, jitter=False, train_frac=0.6, pseed=100),
    # train-mode sampling (one uniform draw per level and ray), HuGS static masks in the data loss (quirk B1)
    'A_train': dict(n=24, seed=22, near=0.2, far=1e6, levels=2, n_prop=64, n_nerf=128, width=256, nerf_depth=8,
                    prop_depth=4, glo=0, contract=True, raydist='reciprocal', opaque=True, jitter=True, train_frac=0.3,
                    pseed=101, transient='withmask'),
    # phototourism_*_withmask.gin shape: no contraction, linear spacing, GLO vectors, 3 levels (repo default 64/64/32)
    'photo': dict(n=20, seed=23, near=1.0, far=2.0, levels=3, n_prop=64, n_nerf=32, width=256, nerf_depth=8,
                  prop_depth=4, glo=4, contract=False, raydist=None, opaque=False, jitter=True, train_frac=1.0,
                  pseed=102, transient='withmask'),
}


def golden_params(c, feat=504, basis_n=21):
  """flax-named parameter tree (NumPy float32) of golden case `c`, regenerated from c['pseed']:
  he_uniform kernels, small non-zero biases (so the bias path is exercised), N(0, 1/sqrt(g)) GLO rows."""
  rng = np.random.default_rng(c['pseed'])
  W, view_in = c['width'], 3 + 6 * 4

  def mlp(depth, rgb, glo):
    shapes, d = [], feat
    for i in range(depth):
      shapes.append((d, W)); d = W
      if i % 4 == 0 and i > 0:
        d = W + feat
    shapes.append((d, 1))
    if rgb:
      shapes += [(d, 256), (256 + view_in + glo, 128), (128, 3)]
    layers = {}
    for i, (fi, fo) in enumerate(shapes):
      bound = np.sqrt(6.0 / fi)
      layers[f'Dense_{i}'] = {'kernel': rng.uniform(-bound, bound, (fi, fo)).astype(np.float32),
                              'bias': rng.uniform(-0.1, 0.1, (fo,)).astype(np.float32)}
    return layers

  tree = {'NerfMLP_0': mlp(c['nerf_depth'], True, c['glo']), 'PropMLP_0': mlp(c['prop_depth'], False, 0)}
  if c['glo'] > 0:
    tree['GloEmbed_0'] = {'embedding': (rng.normal(size=(16, c['glo'])) / np.sqrt(c['glo'])).astype(np.float32)}
  return tree


def golden_jitter(c):
  """Unit-uniform draws [levels][n, 1] handed to stepfun.sample in place of jax.random.uniform."""
  rng = np.random.default_rng(c['pseed'] + 7)
  return [rng.uniform(0, 1, (c['n'], 1)).astype(np.float32) for _ in range(c['levels'])]


def param_checksum(tree):
  def leaves(t):
    if isinstance(t, dict):
      for k in sorted(t):
        yield from leaves(t[k])
    else:
      yield np.asarray(t, np.float64)
  vs = list(leaves(tree))
  return [float(sum(v.sum() for v in vs)), float(sum(np.abs(v).sum() for v in vs))]


def golden_case_configs(c, precision='fp32'):
  """(oracle ModelConfig, EngineConfig kwargs) of a golden case."""
  return config_pair(num_levels=c['levels'], n_prop=c['n_prop'], n_nerf=c['n_nerf'], width=c['width'],
                     nerf_depth=c['nerf_depth'], prop_depth=c['prop_depth'], precision=precision,
                     max_rays=max(c['n'], 128), glo=c['glo'], contract=c['contract'], raydist=c['raydist'],
                     opaque=c['opaque'])
