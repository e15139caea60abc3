"""Operator-level parity: CUDA kernels (through the C ABI) vs the CPU oracle on the same seeded inputs.

Integer work (selected CDF interval) must be bit-exact; fp32 tolerances are stated per test.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='fp32', max_rays=512)
  e = Engine(ecfg, H.basis_np())
  yield e
  e.close()


def _hist(rng, n, nb, zero_width=False):
  t = np.sort(rng.uniform(0, 1, (n, nb + 1)).astype(np.float32), -1)
  t[:, 0], t[:, -1] = 0.0, 1.0
  if zero_width:
    t[:, 3] = t[:, 2]
  w = rng.uniform(0, 1, (n, nb)).astype(np.float32) ** 4
  w /= w.sum(-1, keepdims=True)
  return torch.tensor(t), torch.tensor(w)


@pytest.mark.parametrize('nb,ns', [(64, 64), (190, 128), (5, 10), (17, 31), (1, 64)])
def test_invert_cdf_bit_exact(eng, nb, ns):
  """math.sorted_interp on a caller-provided CDF: indices AND values bit-exact vs the oracle."""
  rng = np.random.default_rng(nb * 1000 + ns)
  n = 300
  t, w = _hist(rng, n, nb, zero_width=nb > 8)
  cw = O.integrate_weights(w)
  u = torch.tensor(np.sort(rng.uniform(0, 1 - 1e-7, (n, ns)).astype(np.float32), -1))
  # adversarial: some queries exactly on CDF knots
  u[:, 0] = cw[:, min(1, nb)] if nb > 1 else u[:, 0]
  u = torch.clamp(u, max=float(np.nextafter(np.float32(1), np.float32(0))))
  ref = O.sorted_interp(u, cw, t)
  ref_idx = O.sorted_interp_index(u, cw)
  out, idx = eng.invert_cdf(t, cw, u)
  assert torch.equal(idx.cpu().long(), ref_idx)
  assert torch.equal(out.cpu(), ref), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize('nb,ns,anneal', [(64, 64, 1.0), (190, 128, 0.37), (17, 31, 1.0)])
def test_sample_intervals_vs_oracle(eng, nb, ns, anneal):
  """stepfun.sample_intervals end to end (softmax + CDF + inversion).  exp/log differ by ulps between
  CPU libm and CUDA, so values agree to 2e-6 and any index flip must be a near-tie on the CDF."""
  rng = np.random.default_rng(7)
  n = 256
  t, w = _hist(rng, n, nb, zero_width=nb > 8)
  logits = torch.where(t[:, 1:] > t[:, :-1], anneal * torch.log(w), torch.tensor(-float('inf')))
  for jitter in (None, torch.tensor(rng.uniform(size=(n, 1)).astype(np.float32))):
    u_base, mj = O.sample_u(ns, True, jitter)
    ref = O.sample_intervals(jitter, t, logits, ns, single_jitter=True, domain=(0., 1.))
    out, idx = eng.sample_intervals(t, logits, u_base, jitter, mj, ns, (0., 1.), want_idx=True)
    cw = O.integrate_weights(torch.softmax(logits, -1))
    # A one-ulp change of the CDF (exp/log differ by ulps between CPU libm and CUDA) moves a sample centre
    # by ~ulp / pdf of its bin, so the tolerance is per sample: e_j = 1e-6 * bin_width / bin_mass, capped at
    # the bin width; interval fenceposts are averages / reflections of adjacent centres.
    o, r = out.cpu().numpy().astype(np.float64), ref.numpy().astype(np.float64)
    ii = idx.cpu().long()
    dt = torch.gather(t[:, 1:] - t[:, :-1], 1, ii).numpy().astype(np.float64)
    dm = torch.gather(cw[:, 1:] - cw[:, :-1], 1, ii).numpy().astype(np.float64)
    e = np.minimum(dt, 1e-6 * dt / np.maximum(dm, 1e-30))
    tol = np.empty_like(o)
    tol[:, 1:-1] = 0.5 * (e[:, 1:] + e[:, :-1])
    tol[:, 0] = 1.5 * e[:, 0] + 0.5 * e[:, 1]
    tol[:, -1] = 1.5 * e[:, -1] + 0.5 * e[:, -2]
    err = np.abs(o - r)
    assert (err <= tol + 1e-6).all(), float((err - tol).max())
    assert (err <= 2e-6).mean() > 0.99
    u = u_base.expand(n, ns) if jitter is None else u_base + jitter * mj
    ref_idx = O.sorted_interp_index(u, cw)
    bad = (idx.cpu().long() != ref_idx)
    if bad.any():
      i = ref_idx[bad]
      rows = torch.nonzero(bad)[:, 0]
      gap = torch.minimum((u[bad] - cw[rows, i]).abs(), (u[bad] - cw[rows, torch.clamp(i + 1, max=nb)]).abs())
      assert float(gap.max()) < 4e-7, f'{int(bad.sum())} index flips, not near-ties: {float(gap.max())}'
    assert float(bad.float().mean()) < 1e-3


def test_sample_intervals_rejects_single_sample(eng):
  """stepfun.py:240-241 raises ValueError for num_samples <= 1."""
  from nerf_hugs_b200._lib import HugsError
  t, w = _hist(np.random.default_rng(0), 4, 8)
  with pytest.raises(HugsError, match='num_samples must be > 1'):
    eng.sample_intervals(t, torch.log(w), torch.zeros(1), None, 0.0, 1, (0., 1.))


def test_sample_intervals_empty_batch(eng):
  out = eng.sample_intervals(torch.zeros(0, 9), torch.zeros(0, 8), torch.linspace(0, .9, 4), None, 0.0, 4, (0., 1.))
  assert out.shape == (0, 5)


@pytest.mark.parametrize('nb', [64, 16, 3])
def test_max_dilate_weights_vs_oracle(eng, nb):
  """stepfun.max_dilate_weights(renormalize=True) + [1:-1] trim: fenceposts bit-exact, weights 1e-6."""
  rng = np.random.default_rng(nb)
  n = 200
  t, w = _hist(rng, n, nb, zero_width=nb > 8)
  w = w * torch.tensor(rng.uniform(0.5, 1.0, (n, 1)).astype(np.float32))   # sums <= 1
  dil = 0.0025 + 0.5 / 64
  tr, wr = O.max_dilate_weights(t, w, dil, domain=(0., 1.), renormalize=True)
  tr, wr = tr[:, 1:-1], wr[:, 1:-1]
  to, wo = eng.max_dilate_weights(t, w, dil, (0., 1.))
  assert torch.equal(to.cpu(), tr)
  np.testing.assert_allclose(wo.cpu().numpy(), wr.numpy(), rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize('S,opaque', [(128, True), (64, False), (32, True), (7, False)])
def test_alpha_composite_vs_oracle(S, opaque):
  """render.compute_alpha_weights + volumetric_rendering incl. extras; rel 2e-5 / abs 2e-6."""
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='fp32', max_rays=64, opaque=opaque)
  e = Engine(ecfg, H.basis_np())
  rng = np.random.default_rng(S)
  n = 200
  tdist = torch.tensor(np.sort(rng.uniform(0.2, 30, (n, S + 1)).astype(np.float32), -1))
  raw_d = torch.tensor(rng.normal(size=(n, S)).astype(np.float32) * 3)
  raw_d[:5] = -50.0                                  # empty rays
  raw_d[5:10, S // 2] = 60.0                         # delta density (render_test.py:443-463)
  raw_rgb = torch.tensor(rng.normal(size=(n, S, 3)).astype(np.float32) * 2)
  dirs = torch.tensor(rng.normal(size=(n, 3)).astype(np.float32))
  far = torch.full((n, 1), 40.0)
  for rgb_in in (raw_rgb, None):
    dens = torch.nn.functional.softplus(raw_d - 1)
    w = O.compute_alpha_weights(dens, tdist, dirs, opaque_background=opaque)[0]
    rgbs = (torch.sigmoid(raw_rgb) * 1.002 - 0.001) if rgb_in is not None else torch.zeros(n, S, 3)
    ref = O.volumetric_rendering(rgbs, w, tdist, 1.0, far, True)
    out = e.alpha_composite(raw_d, rgb_in, tdist, dirs, far)
    np.testing.assert_allclose(out['weights'].cpu().numpy(), w.numpy(), rtol=2e-5, atol=2e-7)
    for k in ('rgb', 'acc', 'distance_mean', 'distance_median', 'distance_percentile_5', 'distance_percentile_95'):
      np.testing.assert_allclose(out[k].cpu().numpy(), ref[k].numpy(), rtol=5e-5, atol=5e-6, err_msg=k)
  e.close()


@pytest.mark.parametrize('contract', [True, False])
def test_ipe_features_vs_oracle(eng, contract):
  """cast_rays -> track_linearize(contract) -> lift_and_diagonalize -> integrated_pos_enc, reference
  column order.  The phase 2^k*mu amplifies one-ulp differences of mu by 2^k, hence the degree-aware bound."""
  n, S = 64, 32
  rays, _ = H.make_rays(n, seed=3)
  rng = np.random.default_rng(4)
  tdist = torch.tensor(np.sort(rng.uniform(0.2, 6.0 if contract else 2.0, (n, S + 1)).astype(np.float32), -1))
  basis = torch.tensor(H.basis_np())
  means, covs = O.cast_rays(tdist, rays['origins'], rays['directions'], rays['radii'], 'cone', diag=False)
  if contract:
    means, covs = O.track_linearize_contract(means, covs)
  lm, lv = O.lift_and_diagonalize(means, covs, basis)
  ref = O.integrated_pos_enc(lm, lv, 0, 12).numpy()
  out = eng.ipe_features(rays, tdist, contract).cpu().numpy()
  assert out.shape == ref.shape == (n, S, 504)
  deg = np.tile(np.repeat(np.arange(12), 21), 2)
  tol = 1e-6 + 4e-7 * (2.0 ** deg) * 3.0
  err = np.abs(out - ref)
  assert (err <= tol).all(), f'max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}'
  # low degrees (where amplification is < 16) are tight
  assert err[..., deg < 4].max() < 2e-5


@pytest.mark.parametrize('name', ['small', 'prop', 'nerf', 'odd'])
@pytest.mark.parametrize('anneal', [1.0, 0.3])
def test_sample_intervals_vs_reference_golden(eng, name, anneal):
  """hugs_sample_intervals directly against outputs of the REFERENCE's own torch twin
  (nerfacto/utils/ray_utils.py:198-223, fixtures tests/golden/sample_intervals.npz), deterministic branch."""
  z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sample_intervals.npz'))
  t = torch.tensor(z[f'{name}_a{anneal}_t'])
  w = torch.tensor(z[f'{name}_a{anneal}_w'])
  ref = z[f'{name}_a{anneal}_out']
  ns = ref.shape[-1] - 1
  logits = torch.where(t[..., 1:] > t[..., :-1], anneal * torch.log(w), torch.tensor(-float('inf')))   # models.py:191-193
  u_base, mj = O.sample_u(ns, True, None)
  out, idx = eng.sample_intervals(t, logits, u_base, None, mj, ns, (0., 1.), want_idx=True)
  # exp/log ulps differ between CUDA and the CPU libm the reference ran on: a one-ulp CDF change moves a sample centre
  # by ~ulp / pdf of its bin (same per-sample bound as test_sample_intervals_vs_oracle); 99 % agree to 2e-6 outright
  cw = O.integrate_weights(torch.softmax(logits, -1))
  ii = idx.cpu().long()
  dt = torch.gather(t[:, 1:] - t[:, :-1], 1, ii).numpy().astype(np.float64)
  dm = torch.gather(cw[:, 1:] - cw[:, :-1], 1, ii).numpy().astype(np.float64)
  e = np.minimum(dt, 2e-6 * dt / np.maximum(dm, 1e-30))
  tol = np.empty(ref.shape, np.float64)
  tol[:, 1:-1] = 0.5 * (e[:, 1:] + e[:, :-1])
  tol[:, 0] = 1.5 * e[:, 0] + 0.5 * e[:, 1]
  tol[:, -1] = 1.5 * e[:, -1] + 0.5 * e[:, -2]
  err = np.abs(out.cpu().numpy().astype(np.float64) - ref.astype(np.float64))
  assert (err <= tol + 2e-6).all(), float((err - tol).max())
  assert (err <= 2e-6).mean() > 0.99
  assert (np.diff(out.cpu().numpy(), axis=-1) >= 0).all()


def test_sample_intervals_single_interval_known_answer(eng):
  """The reference's RNG-free known answer (stepfun_test.py:579-586): all mass in one interval -> linspace over it."""
  t = torch.tensor([[1., 2, 3, 4, 5, 6]])
  logits = torch.tensor([[0., 0, 100, 0, 0]])
  u_base, mj = O.sample_u(10, True, None)
  out, idx = eng.sample_intervals(t, logits, u_base, None, mj, 10, (0., 10.), want_idx=True)
  np.testing.assert_allclose(out.cpu().numpy()[0], np.linspace(3, 4, 11), atol=1e-5, rtol=1e-5)
  assert (idx.cpu().numpy() == 2).all()


def test_alpha_weights_delta_density_known_answer():
  """render_test.py:443-463: one interval with a huge density gives one-hot weights (atol 1e-5)."""
  from nerf_hugs_b200.engine import Engine
  _, ecfg = H.config_pair(precision='fp32', max_rays=128, opaque=False)
  e = Engine(ecfg, H.basis_np())
  rng = np.random.default_rng(0)
  n, d = 100, 128
  r = rng.normal(size=(n, d))
  mask = (r == r.max(-1, keepdims=True)).astype(np.float32)
  raw_d = torch.tensor(np.where(mask > 0, 1e10, -1e4).astype(np.float32))      # softplus(raw - 1): 1e10 / 0
  tdist = torch.tensor(np.sort(rng.uniform(0.5, 2.5, (n, d + 1)).astype(np.float32), -1))
  dirs = torch.tensor(rng.normal(size=(n, 3)).astype(np.float32))
  out = e.alpha_composite(raw_d, None, tdist, dirs, torch.full((n, 1), 3.0))
  np.testing.assert_allclose(out['weights'].cpu().numpy(), mask, atol=1e-5, rtol=1e-5)
  np.testing.assert_allclose(out['acc'].cpu().numpy(), 1.0, atol=1e-5)
  e.close()
