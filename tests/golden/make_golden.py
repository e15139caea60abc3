"""Generate golden fixtures by running the REFERENCE's own importable code.

Run once in the build container (`python tests/golden/make_golden.py`); the
reference tree (/root/reference) does not exist on the GPU box, so the outputs
are committed as `tests/golden/*.npz` next to this script.

What is importable from the reference without JAX:
  * MipNeRF360/internal/geopoly.py        (pure NumPy)   -> IPE basis, column order
  * nerfacto/utils/ray_utils.py           (torch)        -> sample / sample_intervals /
                                                            density_to_weight (torch twins of
                                                            stepfun.sample*, render.compute_alpha_weights)
  * nerfacto/utils/loss_utils.py          (torch)        -> lossfun_outer / lossfun_distortion
  * nerfacto/models/custom_functions.py   (torch)        -> contraction, pos_enc
  * MipNeRF360/internal/camera_utils.py   (NumPy branch, xnp=np; its jax / internal.* imports are satisfied by empty
                                           stub modules because pixels_to_rays never touches them with xnp=np)
                                                         -> pixels_to_rays, get_pixtocam
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def load_camera_utils():
  class Stub(types.ModuleType):
    def __getattr__(self, k):
      if k.startswith('__'):
        raise AttributeError(k)
      return type(k, (), {})
  saved = {}
  names = ['jax', 'jax.numpy', 'internal', 'internal.configs', 'internal.math', 'internal.stepfun', 'internal.utils']
  for n in names:
    saved[n] = sys.modules.get(n)
    sys.modules[n] = Stub(n)
  sys.modules['jax'].numpy = sys.modules['jax.numpy']
  for n in ('configs', 'math', 'stepfun', 'utils'):
    setattr(sys.modules['internal'], n, sys.modules['internal.' + n])
  try:
    return load(f'{REF}/MipNeRF360/internal/camera_utils.py', 'ref_camera_utils')
  finally:
    for n in names:
      if saved[n] is None:
        sys.modules.pop(n, None)
      else:
        sys.modules[n] = saved[n]


def camera_golden(rng):
  """Rays of random pixels of random cameras through the reference's pixels_to_rays (NumPy branch)."""
  cu = load_camera_utils()
  n_cams, n = 6, 257
  hw = rng.integers(40, 90, size=(n_cams, 2))
  focal = rng.uniform(50., 120., size=n_cams)
  pixtocams = np.stack([cu.get_pixtocam(focal[i], int(hw[i, 1]), int(hw[i, 0])) for i in range(n_cams)]).astype(np.float32)
  rot, _ = np.linalg.qr(rng.normal(size=(n_cams, 3, 3)))
  pos = rng.normal(size=(n_cams, 3, 1)) * 2.0
  camtoworlds = np.concatenate([rot, pos], -1).astype(np.float32)
  cam_idx = rng.integers(0, n_cams, size=n)
  px = (rng.uniform(size=n) * hw[cam_idx, 1]).astype(np.int64)
  py = (rng.uniform(size=n) * hw[cam_idx, 0]).astype(np.int64)
  o, d, v, r = cu.pixels_to_rays(px, py, pixtocams[cam_idx], camtoworlds[cam_idx], xnp=np)
  out = dict(heights=hw[:, 0], widths=hw[:, 1], pixtocams=pixtocams, camtoworlds=camtoworlds,
             cam_idx=cam_idx, pix_x=px, pix_y=py, origins=o, directions=d, viewdirs=v, radii=r)
  # lens distortion (10 Newton iterations, camera_utils.py:460-494) and fisheye projection (:557-568)
  dist = dict(k1=0.12, k2=-0.05, k3=0.01, k4=-0.002, p1=0.003, p2=-0.002)
  out['dist_params'] = np.array([dist[k] for k in ('k1', 'k2', 'k3', 'k4', 'p1', 'p2')], np.float32)
  dist = {k: float(np.float32(v)) for k, v in dist.items()}
  for tag, kw in (('dist', dict(distortion_params=dist)), ('fish', dict(camtype=cu.ProjectionType.FISHEYE)),
                  ('distfish', dict(distortion_params=dist, camtype=cu.ProjectionType.FISHEYE))):
    o2, d2, v2, r2 = cu.pixels_to_rays(px, py, pixtocams[cam_idx], camtoworlds[cam_idx], xnp=np, **kw)
    out.update({f'{tag}_directions': d2, f'{tag}_viewdirs': v2, f'{tag}_radii': r2})
  np.savez(f'{OUT}/camera.npz', **out)


def nerfacto_golden(rng, ray_utils):
  """density_to_weight (quirks B2 / B6), render_features, render_depth (B7), pdf_sample of nerfacto/utils/ray_utils.py."""
  n, S = 48, 64
  bins = np.sort(rng.uniform(0.5, 6.0, (n, S + 1)).astype(np.float32), -1)
  dens = (rng.uniform(0, 1, (n, S)).astype(np.float32) ** 3) * 8
  dens[:3] = 0.0
  dirs = rng.normal(size=(n, 3)).astype(np.float32)
  feats = rng.uniform(size=(n, S, 3)).astype(np.float32)
  bg = rng.uniform(size=(n, 3)).astype(np.float32)
  out = {'bins': bins, 'dens': dens, 'dirs': dirs, 'feats': feats, 'bg': bg}
  for opaque in (False, True):
    w, a, tr = ray_utils.density_to_weight(torch.tensor(dens), torch.tensor(bins), torch.tensor(dirs), opaque)
    out[f'w_{int(opaque)}'], out[f'a_{int(opaque)}'], out[f't_{int(opaque)}'] = w.numpy(), a.numpy(), tr.numpy()
    out[f'rgb_{int(opaque)}'] = ray_utils.render_features(w, torch.tensor(feats), torch.tensor(bg), False).numpy()
    out[f'depth_{int(opaque)}'] = ray_utils.render_depth(w, torch.tensor(bins)).numpy()
  sb = np.sort(rng.uniform(0, 1, (n, S + 1)).astype(np.float32), -1); sb[:, 0], sb[:, -1] = 0.0, 1.0
  ww = rng.uniform(0, 1, (n, S)).astype(np.float32) ** 4
  ww[:2] = 0.0
  out['pdf_bins'], out['pdf_w'] = sb, ww
  out['pdf_out'] = ray_utils.pdf_sample(torch.tensor(sb), torch.tensor(ww), 32, False, True).numpy()
  out['uniform_out'] = ray_utils.uniform_sample(5, 16, False, True, 'cpu').numpy()
  np.savez(f'{OUT}/nerfacto_ops.npz', **out)


def main():
  torch.manual_seed(0)
  rng = np.random.default_rng(0)

  camera_golden(np.random.default_rng(7))

  geopoly = load(f'{REF}/MipNeRF360/internal/geopoly.py', 'ref_geopoly')
  np.savez(f'{OUT}/geopoly_basis.npz',
           icosahedron_2=geopoly.generate_basis('icosahedron', 2),
           octahedron_1=geopoly.generate_basis('octahedron', 1),
           octahedron_4=geopoly.generate_basis('octahedron', 4),
           icosahedron_1=geopoly.generate_basis('icosahedron', 1))

  ray_utils = load(f'{REF}/nerfacto/utils/ray_utils.py', 'ref_ray_utils')
  loss_utils = load(f'{REF}/nerfacto/utils/loss_utils.py', 'ref_loss_utils')
  cf = load(f'{REF}/nerfacto/models/custom_functions.py', 'ref_custom_functions')

  nerfacto_golden(np.random.default_rng(11), ray_utils)

  # --- sample_intervals (deterministic branch == stepfun.sample_intervals(rng=None)) ---
  cases = {}
  for name, (n_rays, n_bins, n_samples) in {
      'small': (7, 5, 10), 'prop': (64, 64, 64), 'nerf': (64, 190, 128), 'odd': (33, 17, 31)}.items():
    t = np.sort(rng.uniform(0, 1, (n_rays, n_bins + 1)).astype(np.float32), -1)
    t[:, 0], t[:, -1] = 0.0, 1.0
    if name == 'odd':           # zero-width bins (weight forced to -inf by the reference)
      t[:, 5] = t[:, 4]
    w = rng.uniform(0, 1, (n_rays, n_bins)).astype(np.float32) ** 4
    w /= w.sum(-1, keepdims=True)
    for anneal in (1.0, 0.3):
      out = ray_utils.sample_intervals(torch.tensor(t), torch.tensor(w), anneal, 0.0, n_samples,
                                       perturb=False, single_jitter=True, domain=(0.0, 1.0))
      cases[f'{name}_a{anneal}_t'] = t
      cases[f'{name}_a{anneal}_w'] = w
      cases[f'{name}_a{anneal}_out'] = out.numpy()
  # the reference's own known-answer case (stepfun_test.py:579-586)
  t = torch.tensor([[1., 2, 3, 4, 5, 6]])
  w = torch.softmax(torch.tensor([[0., 0, 100, 0, 0]]), -1)
  cases['single_out'] = ray_utils.sample_intervals(t, w, 1.0, 0.0, 10, False, True,
                                                   (-float('inf'), float('inf'))).numpy()
  np.savez(f'{OUT}/sample_intervals.npz', **cases)

  # --- lossfun_distortion / lossfun_outer ---
  n_rays = 32
  t = np.sort(rng.uniform(0, 1, (n_rays, 129)).astype(np.float32), -1)
  w = rng.uniform(0, 1, (n_rays, 128)).astype(np.float32); w /= (1.3 * w.sum(-1, keepdims=True))
  te = np.sort(rng.uniform(0, 1, (n_rays, 65)).astype(np.float32), -1)
  te[:, 0], te[:, -1] = 0.0, 1.0
  t[:, 0] = np.maximum(t[:, 0], 0.0)
  we = rng.uniform(0, 1, (n_rays, 64)).astype(np.float32); we /= (1.1 * we.sum(-1, keepdims=True))
  np.savez(f'{OUT}/losses.npz', t=t, w=w, t_env=te, w_env=we,
           distortion=loss_utils.lossfun_distortion(torch.tensor(t), torch.tensor(w)).numpy(),
           outer=loss_utils.outer(torch.tensor(t[:, :-1]), torch.tensor(t[:, 1:]),
                                  torch.tensor(te[:, :-1]), torch.tensor(te[:, 1:]),
                                  torch.tensor(we)).numpy())

  # --- contraction + pos_enc ---
  x = (rng.normal(size=(256, 3)) * np.array([0.3, 2.0, 30.0])).astype(np.float32)
  v = rng.normal(size=(64, 3)).astype(np.float32)
  v /= np.linalg.norm(v, axis=-1, keepdims=True)
  np.savez(f'{OUT}/coord.npz', x=x, contract=cf.spatial_distortion_norm2(torch.tensor(x)).numpy(),
           v=v, pos_enc_0_4=cf.pos_enc(torch.tensor(v), 0, 4, True).numpy())

  print('wrote fixtures to', OUT)


if __name__ == '__main__':
  main()
