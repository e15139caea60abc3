"""CUDA operator kernels (through the C ABI) against outputs of the reference's OWN Mip-NeRF 360 source.

Fixtures: tests/golden/mip360_ops.npz, produced by tests/golden/make_golden_mipnerf360.py, which executes
/root/reference/MipNeRF360/internal/{math,stepfun,coord,render}.py on a NumPy stand-in for jax (float32).
fp32 values: 1e-6 absolute unless a comment states why not; selected CDF intervals exact.
"""
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def ops():
  return np.load(f'{G}/mip360_ops.npz')


@pytest.fixture(scope='module')
def eng():
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='fp32', max_rays=512, opaque=False)
  e = Engine(ecfg, H.basis_np())
  yield e
  e.close()


def T(x):
  return torch.tensor(np.asarray(x))


def test_sorted_interp_bit_exact(ops, eng):
  """math.sorted_interp (math.py:108-127): hugs_invert_cdf(t=fp, cw=xp, u=x) equals the reference bit for bit."""
  out, _ = eng.invert_cdf(T(ops['si_fp']), T(ops['si_xp']), T(ops['si_x']))
  assert np.array_equal(out.cpu().numpy(), ops['si_out'])


@pytest.mark.parametrize('dil', ['0.0103', '0.0200'])
def test_max_dilate_weights(ops, eng, dil):
  """stepfun.max_dilate_weights (stepfun.py:89-128) + the [1:-1] trim of models.py:178-179."""
  d = float(np.float32(0.02 if dil == '0.0200' else 0.0025 + 0.5 / 64))
  td, wd = eng.max_dilate_weights(T(ops['dilate_in_t']), T(ops['dilate_in_w']), d, (0., 1.))
  assert np.array_equal(td.cpu().numpy(), ops[f'dilate_{dil}_t'][:, 1:-1])          # fenceposts: bit-exact
  np.testing.assert_allclose(wd.cpu().numpy(), ops[f'dilate_{dil}_w'][:, 1:-1], atol=1e-7, rtol=2e-6)


def test_sample_intervals_det_and_jitter(ops, eng):
  """stepfun.sample_intervals (stepfun.py:214-263), rng=None and single-jitter branches."""
  from oracle import mipnerf360 as O
  t, logits = T(ops['dilate_in_t']), T(ops['samp_logits'])
  u_det, _ = O.sample_u(32, True, None)
  out = eng.sample_intervals(t, logits, u_det, None, 0.0, 32, (0., 1.))
  np.testing.assert_allclose(out.cpu().numpy(), ops['samp_det'], atol=3e-6)      # exp / log ulps, cumsum order
  jit = T(ops['samp_jitter_u'])
  u_tr, max_jitter = O.sample_u(32, True, jit)
  out = eng.sample_intervals(t, logits, u_tr, jit[:, 0], max_jitter, 32, (0., 1.))
  np.testing.assert_allclose(out.cpu().numpy(), ops['samp_jitter'], atol=3e-6)


@pytest.mark.parametrize('opaque', [0, 1])
def test_alpha_composite(ops, opaque):
  """render.compute_alpha_weights + volumetric_rendering with extras (render.py:130-151,185-244)."""
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='fp32', max_rays=64, opaque=bool(opaque))
  e = Engine(ecfg, H.basis_np())
  dens, rgbs = ops['vr_density'], ops['vr_rgbs']
  # the ABI takes pre-activation values: invert softplus(raw - 1) and sigmoid(raw) * 1.002 - 0.001 in float64
  d64 = dens.astype(np.float64)
  raw_d = np.where(d64 > 30, d64, np.log(np.expm1(np.maximum(d64, 1e-30)))) + 1.0
  raw_d = np.where(d64 == 0, -80.0, raw_d).astype(np.float32)
  c = (rgbs.astype(np.float64) + 0.001) / 1.002
  raw_c = np.log(c / (1 - c)).astype(np.float32)
  out = e.alpha_composite(T(raw_d), T(raw_c), T(ops['vr_tdist']), T(ops['vr_dirs']), T(ops['vr_far']))
  # weights go through softplus(inverse softplus) in float32: 2e-6 relative
  np.testing.assert_allclose(out['weights'].cpu().numpy(), ops[f'vr_w_{opaque}'], rtol=5e-6, atol=3e-7)
  for k in ('rgb', 'acc'):
    np.testing.assert_allclose(out[k].cpu().numpy(), ops[f'vr_{k}_{opaque}'], atol=2e-6, err_msg=k)
  for k in ('distance_mean', 'distance_median', 'distance_percentile_5', 'distance_percentile_95'):
    np.testing.assert_allclose(out[k].cpu().numpy(), ops[f'vr_{k}_{opaque}'], rtol=2e-5, err_msg=k)
  e.close()


@pytest.mark.parametrize('contract', [True, False])
def test_ipe_features_exact_path(ops, eng, contract):
  """cast_rays -> track_linearize(contract) -> lift_and_diagonalize -> integrated_pos_enc incl. safe_sin (B12) at
  phases up to ~4096 rad.  The lifted mean differs from the reference's by fp32 rounding, amplified by 2^k."""
  rays = dict(origins=T(ops['cast_o']), directions=T(ops['cast_d']), viewdirs=T(ops['cast_d']), radii=T(ops['cast_radii']),
              near=torch.ones(12, 1), far=torch.ones(12, 1))
  out = eng.ipe_features(rays, T(ops['cast_t']), contract).cpu().numpy()
  ref = ops['ipe_contract'] if contract else ops['ipe_plain']
  m = np.ones(ref.shape[:2], bool) if contract else ops['ipe_plain_mask']
  deg = np.tile(np.repeat(np.arange(12), 21), 2)
  err = np.abs(out - ref)[m]
  # contracted coordinates are <= 2: one ulp of the mean is 2.4e-7, times 2^k radians of phase; the fixture's far
  # samples (t up to 3000) add a few ulps through o + d t and the contraction itself
  tol = 2e-6 + 5e-6 * (2.0 ** deg)
  assert (err <= tol).all(), f'max err {err.max()}'
  assert err[:, deg < 3].max() < 2e-5


def test_bf16_encoder_op_level(ops):
  """The throughput-mode encoder (encode_bf16_kernel: phase reduction in turns + sin.approx + angle doubling) on its own:
  bf16-rounded features within bf16 resolution (2^-8 relative, 4e-3 absolute) of the reference's features, and without
  any systematic phase drift beyond the reference's own B12 wrap (8e-5 rad)."""
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='bf16_tc', max_rays=128)
  e = Engine(ecfg, H.basis_np())
  rays = dict(origins=T(ops['cast_o']), directions=T(ops['cast_d']), viewdirs=T(ops['cast_d']), radii=T(ops['cast_radii']),
              near=torch.ones(12, 1), far=torch.ones(12, 1))
  feat = e.debug_encode_bf16(rays, T(ops['cast_t']), True).float().cpu().numpy()     # [n*S, 512] engine order
  ref = ops['ipe_contract'].reshape(-1, 504)
  # engine column f' = (b * 12 + k) * 2 + s  <->  reference column s * 252 + k * 21 + b
  fp = np.arange(504)
  s, bk = fp & 1, fp >> 1
  b, k = bk // 12, bk % 12
  got = feat[:, :504]
  want = ref[:, s * 252 + k * 21 + b]
  assert np.all(feat[:, 504:] == 0)
  err = np.abs(got - want)
  tol = 4.5e-3 + 5e-6 * (2.0 ** k)[None, :]
  assert (err <= tol).all(), f'max err {err.max()} (degree {k[np.unravel_index(err.argmax(), err.shape)[1]]})'
  assert float(np.abs((got - want)[:, k < 6]).mean()) < 1.2e-3     # mean rounding error of bf16 (~2^-10 of |x| <= 1)
  e.close()
