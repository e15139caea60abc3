"""Full-size (BASELINE.json config A: 4096 rays x (64 + 128) samples) checks through size-independent properties.

The CPU oracle cannot run 4096 rays in seconds, so the full batch is checked through properties of the path:

* rays are independent: ray i of a 4096-ray batch renders exactly like ray i of a small batch (tensor-core tiles, the
  CTA-pair scheduling and padding tiles must not leak between rows), and a slice of the batch matches the oracle;
* the gradient is linear in the per-ray loss weights (`lossmult`): zero-weight rays contribute nothing to the data term;
* a batch whose tile count is not a multiple of the 4-tile CTA-pair unit gives the same per-ray results.
"""
import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O
from tests import helpers as H
from tests.test_gpu_train import _loss_cfg

pytestmark = pytest.mark.gpu

N_FULL = 4096


def _engine(n, precision='bf16_tc'):
  from nerf_hugs_b200.engine import Engine
  ocfg, ecfg = H.config_pair(precision=precision, max_rays=n)
  params = O.init_params(ocfg, seed=0, bias_scale=0.1)
  eng = Engine(ecfg, H.basis_np())
  flat = eng.flatten_params(params)
  eng.params_changed(flat)
  return ocfg, params, eng, flat


def _slice(rays, sl):
  return {k: v[sl] for k, v in rays.items()}


def test_full_batch_rays_are_independent_and_match_oracle():
  ocfg, params, eng, flat = _engine(N_FULL)
  rays, _ = H.make_rays(N_FULL, seed=3)
  res, _ = eng.forward(flat, rays, 0.5, None, compute_extras=True)
  full = {k: v.cpu() for k, v in res[-1].items()}
  # (a) the same rays in small batches: bit-identical per-ray results (48 rays: 1.5 units of tiles, ragged tail)
  for sl in (slice(0, 48), slice(1000, 1037), slice(N_FULL - 5, N_FULL)):
    sub, _ = eng.forward(flat, _slice(rays, sl), 0.5, None, compute_extras=True)
    for k in ('rgb', 'acc', 'distance_mean', 'distance_median'):
      assert torch.equal(sub[-1][k].cpu(), full[k][sl]), (k, sl)
  # (b) a slice against the CPU oracle run with bf16-rounded Dense operands (tolerances of test_gpu_model.py)
  sl = slice(2048, 2048 + 32)
  with torch.no_grad():
    rend, _ = O.model_apply(ocfg, params, _slice(rays, sl), 0.5, True, torch.tensor(H.basis_np()), jitter=None, quant='bf16')
  err = float((full['rgb'][sl] - rend[-1]['rgb']).abs().max() / rend[-1]['rgb'].abs().max())
  assert err < 2e-3, err
  assert torch.isfinite(full['rgb']).all() and torch.isfinite(full['distance_mean']).all()
  eng.close()


def test_full_batch_gradient_is_linear_in_lossmult():
  """The data term is sum_i lossmult_i * loss_i / sum_i lossmult_i: with the interlevel / distortion terms switched
  off, zeroing the weights of half of the rays must give the gradient of the other half alone."""
  ocfg, params, eng, flat = _engine(N_FULL)
  rays, gt = H.make_rays(N_FULL, seed=4)
  lcfg = O.LossConfig(distortion_loss_mult=0.0, interlevel_loss_mult=0.0)
  jit = torch.rand(2, N_FULL, generator=torch.Generator().manual_seed(7))
  half = N_FULL // 2
  r_half = {k: v.clone() for k, v in rays.items()}
  r_half['lossmult'][half:] = 0.0
  g_masked, st_masked = eng.loss_and_grad(flat, r_half, gt, 0.5, jit, _loss_cfg(lcfg))
  g_masked = g_masked.clone(); st_masked = st_masked.clone()
  g_sub, st_sub = eng.loss_and_grad(flat, _slice(rays, slice(0, half)), gt[:half], 0.5, jit[:, :half].contiguous(), _loss_cfg(lcfg))
  torch.cuda.synchronize()
  assert torch.isfinite(g_masked).all() and torch.isfinite(g_sub).all()
  np.testing.assert_allclose(float(st_masked[1]), float(st_sub[1]), rtol=1e-5)
  # fp32 atomics reorder the sample reduction between the two runs: compare norms of the difference
  rel = float((g_masked - g_sub).norm() / g_sub.norm())
  assert rel < 1e-3, rel
  eng.close()


@pytest.mark.parametrize('n', [1, 3, 130])
def test_tile_counts_off_the_unit_boundary(n):
  """n rays -> n NeRF tiles / n/2 proposal tiles: every remainder of the 4-tile CTA-pair unit, incl. a half tile."""
  ocfg, params, eng, flat = _engine(256)
  rays, gt = H.make_rays(256, seed=5)
  ref, _ = eng.forward(flat, rays, 0.5, None, compute_extras=True)
  sub, _ = eng.forward(flat, _slice(rays, slice(0, n)), 0.5, None, compute_extras=True)
  for k in ('rgb', 'acc', 'distance_mean'):
    assert torch.equal(sub[-1][k].cpu(), ref[-1][k][:n].cpu()), k
  lcfg = O.LossConfig()
  jit = torch.rand(2, n, generator=torch.Generator().manual_seed(9))
  g, st = eng.loss_and_grad(flat, _slice(rays, slice(0, n)), gt[:n], 0.5, jit, _loss_cfg(lcfg))
  torch.cuda.synchronize()
  assert torch.isfinite(g).all() and torch.isfinite(st).all() and float(g.norm()) > 0
  eng.close()
