"""Re-assert, on the CPU oracle, the properties the reference's unit tests assert.

Each test names the reference test it transfers (MipNeRF360/tests/*).  The
reference's tests are seeded with jax.random, which is not available here, so
the *property* is re-checked with NumPy seeds.
"""
import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O


def test_searchsorted_bounds():
  """stepfun_test.py:108-124 (in range, vs np.searchsorted)."""
  rng = np.random.default_rng(0)
  a = np.sort(rng.uniform(size=(10, 50)).astype(np.float32), -1)
  v = rng.uniform(a[:, :1], a[:, -1:], size=(10, 30)).astype(np.float32)
  lo, hi = O.searchsorted(torch.tensor(a), torch.tensor(v))
  for i in range(10):
    ref = np.searchsorted(a[i], v[i], side='right')
    np.testing.assert_array_equal(hi[i].numpy(), ref)
    np.testing.assert_array_equal(lo[i].numpy(), ref - 1)


def test_searchsorted_out_of_range():
  """stepfun_test.py:55-106."""
  a = torch.linspace(0, 1, 11)[None]
  lo, hi = O.searchsorted(a, torch.tensor([[-1.0, 2.0]]))
  assert lo.tolist() == [[0, 10]] and hi.tolist() == [[0, 10]]


def test_sample_intervals_unbiased_deterministic():
  """stepfun_test.py:542-565 (deterministic, bounded and unbounded)."""
  t = torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5]).repeat(20, 1)
  logits = torch.tensor([0, 0, 100., 0, 0]).repeat(20, 1)
  for domain in ((-0.5, 0.5), (-float('inf'), float('inf'))):
    ts = O.sample_intervals(None, t, logits, 64, single_jitter=True, domain=domain)
    np.testing.assert_allclose(ts.mean(-1).numpy(), 0, atol=1e-5)


def test_sample_train_pdf_reproduced():
  """stepfun_test.py:305-383 (property): jittered samples reproduce the PDF."""
  rng = np.random.default_rng(1)
  nb, ns, nr = 8, 256, 4000
  t = np.sort(rng.uniform(size=nb + 1)).astype(np.float32)
  logits = rng.normal(size=nb).astype(np.float32)
  jit = torch.tensor(rng.uniform(size=(nr, 1)).astype(np.float32))
  s = O.sample(jit, torch.tensor(t).repeat(nr, 1), torch.tensor(logits).repeat(nr, 1), ns,
               single_jitter=True)
  hist = np.histogram(s.numpy().ravel(), bins=t)[0] / (nr * ns)
  p = np.exp(logits) / np.exp(logits).sum()
  np.testing.assert_allclose(hist, p, atol=2e-3)


def test_inner_outer_vs_loops():
  """stepfun_test.py:27-50,699-737."""
  rng = np.random.default_rng(2)
  t0 = np.sort(rng.uniform(size=17)); t1 = np.sort(rng.uniform(size=9)); y1 = rng.uniform(size=8)
  inner, outer = O.inner_outer(torch.tensor(t0)[None], torch.tensor(t1)[None], torch.tensor(y1)[None])
  for i in range(16):
    lo, hi = t0[i], t0[i + 1]
    if lo < t1[0] or hi > t1[-1]:
      continue
    o = sum(y1[j] for j in range(8) if t1[j] < hi and t1[j + 1] > lo)
    inn = sum(y1[j] for j in range(8) if t1[j] >= lo and t1[j + 1] <= hi)
    np.testing.assert_allclose(outer[0, i].item(), o, atol=1e-12)
    np.testing.assert_allclose(inner[0, i].item(), inn, atol=1e-12)


def test_lossfun_outer_same_histogram_is_zero():
  """stepfun_test.py:588-610 ('sameset')."""
  rng = np.random.default_rng(3)
  t = torch.tensor(np.sort(rng.uniform(size=(4, 33))))
  w = torch.tensor(rng.uniform(size=(4, 32)))
  assert float(O.lossfun_outer(t, w, t, w).abs().max()) < 1e-12


def test_distortion_prefix_sum_identity():
  """stepfun_test.py:252-273 (property) + SURVEY App. A: O(S) form == O(S^2) definition."""
  rng = np.random.default_rng(4)
  t = np.sort(rng.uniform(size=(5, 129))); w = rng.uniform(size=(5, 128)) / 128
  ref = O.lossfun_distortion(torch.tensor(t), torch.tensor(w)).numpy()
  u = 0.5 * (t[:, 1:] + t[:, :-1])
  W = np.cumsum(w, -1) - w
  WU = np.cumsum(w * u, -1) - w * u
  fast = 2 * np.sum(w * (u * W - WU), -1) + np.sum(w ** 2 * np.diff(t), -1) / 3
  np.testing.assert_allclose(fast, ref, rtol=1e-10)


def test_max_dilate_vs_brute_force_queries():
  """stepfun_test.py:275-300."""
  rng = np.random.default_rng(5)
  t = np.sort(rng.uniform(size=33)); w = rng.uniform(size=32); d = 0.02
  td, wd = O.max_dilate(torch.tensor(t)[None], torch.tensor(w)[None], d)
  td, wd = td[0].numpy(), wd[0].numpy()
  for q in rng.uniform(td[0], td[-1], size=200):
    i = np.searchsorted(td, q, side='right') - 1
    if i < 0 or i >= len(wd):
      continue
    ref = max([w[j] for j in range(32) if t[j] - d <= q < t[j + 1] + d] + [0.0])
    # the dilated step function must upper-bound (and here equal) the brute-force max
    assert wd[i] >= ref - 1e-12


def test_weighted_percentile():
  """stepfun_test.py:739-790 (property): percentiles of a uniform histogram are linear."""
  t = torch.linspace(0, 1, 11, dtype=torch.float64)[None]
  w = torch.full((1, 10), 0.1, dtype=torch.float64)
  out = O.weighted_percentile(t, w, [5, 50, 95])
  np.testing.assert_allclose(out.numpy()[0], [0.05, 0.5, 0.95], atol=1e-12)


def test_contract_properties_and_jacobian():
  """coord_test.py:61-110 + closed-form Jacobian vs autograd (replaces jax.linearize)."""
  rng = np.random.default_rng(6)
  x = torch.tensor(rng.normal(size=(200, 3)) * 3, dtype=torch.float64)
  z = O.contract(x)
  assert float(z.norm(dim=-1).max()) < 2.0
  inside = x.norm(dim=-1) <= 1
  assert torch.equal(z[inside], x[inside])
  for i in range(0, 200, 17):
    J = torch.autograd.functional.jacobian(O.contract, x[i])
    v = torch.tensor(rng.normal(size=3))
    np.testing.assert_allclose(O.contract_jacobian_apply(x[i], v).numpy(), (J @ v).numpy(), atol=1e-10)


def test_contract_reciprocal_warp_equal_steps():
  """coord_test.py:61-69 style: s_to_t/t_to_s closed form (coord_test.py:199-221)."""
  near, far = torch.tensor([[0.2]]), torch.tensor([[1e6]])
  t_to_s, s_to_t = O.construct_ray_warps('reciprocal', near, far)
  s = torch.linspace(0, 1, 9)[None]
  t = s_to_t(s)
  np.testing.assert_allclose(t_to_s(t).numpy(), s.numpy(), atol=1e-6)
  np.testing.assert_allclose(t[0, 0].item(), 0.2, rtol=1e-6)
  np.testing.assert_allclose(t[0, -1].item(), 1e6, rtol=1e-6)


def test_ipe_zero_variance_is_pe():
  """coord_test.py:129-140."""
  rng = np.random.default_rng(7)
  x = torch.tensor(rng.uniform(-1, 1, size=(20, 3)).astype(np.float32))
  ipe = O.integrated_pos_enc(x, torch.zeros_like(x), 0, 5)
  pe = O.pos_enc(x, 0, 5, append_identity=False)
  np.testing.assert_allclose(ipe.numpy(), pe.numpy(), atol=1e-6)


def test_ipe_vs_monte_carlo():
  """coord_test.py:230-261 (property)."""
  rng = np.random.default_rng(8)
  mean = rng.uniform(-1, 1, size=3); var = rng.uniform(0.01, 0.1, size=3)
  s = rng.normal(size=(200000, 3)) * np.sqrt(var) + mean
  mc = O.pos_enc(torch.tensor(s), 0, 3, append_identity=False).mean(0).numpy()
  ipe = O.integrated_pos_enc(torch.tensor(mean)[None], torch.tensor(var)[None], 0, 3)[0].numpy()
  np.testing.assert_allclose(ipe, mc, atol=5e-3)


def test_safe_sin_large_inputs():
  """math_test.py:38-47: |safe_sin - sin| < 1e-4 ... holds below the wrap; finite above."""
  x = torch.linspace(-300, 300, 1001)
  np.testing.assert_allclose(O.safe_sin(x).numpy(), np.sin(x.numpy().astype(np.float64)), atol=1e-4)
  assert torch.isfinite(O.safe_sin(torch.tensor([1e10, -1e10]))).all()


def test_alpha_weights_delta_density_is_one_hot():
  """render_test.py:443-463."""
  n = 16
  tdist = torch.linspace(1, 2, n + 1)[None]
  density = torch.zeros(1, n); density[0, 5] = 1e10
  w = O.compute_alpha_weights(density, tdist, torch.tensor([[0., 0, 1]]))[0]
  ref = torch.zeros(1, n); ref[0, 5] = 1
  np.testing.assert_allclose(w.numpy(), ref.numpy(), atol=1e-6)


def test_alpha_weights_finite_over_magnitudes():
  """render_test.py:408-441 (value part)."""
  for dm in (1e-8, 1e-2, 1e2, 1e8):
    for tm in (1e-4, 1, 1e4):
      tdist = torch.linspace(0, 1, 33)[None] * tm
      w, a, tr = O.compute_alpha_weights(torch.full((1, 32), dm), tdist, torch.tensor([[1., 0, 0]]),
                                         opaque_background=True)
      assert torch.isfinite(w).all() and abs(float(w.sum()) - 1) < 1e-5


def test_conical_frustum_vs_monte_carlo():
  """render_test.py:279-318 (property)."""
  rng = np.random.default_rng(9)
  d = np.array([0.3, -0.5, 0.8]); t0, t1, r = 1.0, 1.6, 0.1
  mean, cov = O.conical_frustum_to_gaussian(torch.tensor(d)[None], torch.tensor([[t0]]),
                                            torch.tensor([[t1]]), torch.tensor([[r]]), diag=False)
  # sample the frustum: t ~ p(t) ∝ t^2, disc radius r*t*|d|... perpendicular to d
  n = 400000
  tt = (rng.uniform(size=n) * (t1 ** 3 - t0 ** 3) + t0 ** 3) ** (1 / 3)
  rad = r * tt * np.sqrt(rng.uniform(size=n)); th = rng.uniform(0, 2 * np.pi, n)
  dn = d / np.linalg.norm(d)
  a = np.cross(dn, [1, 0, 0]); a /= np.linalg.norm(a); b = np.cross(dn, a)
  pts = tt[:, None] * d + (rad * np.cos(th))[:, None] * a * np.linalg.norm(d) + (rad * np.sin(th))[:, None] * b * np.linalg.norm(d)
  np.testing.assert_allclose(mean[0, 0].numpy(), pts.mean(0), atol=3e-3)
  np.testing.assert_allclose(cov[0, 0].numpy(), np.cov(pts.T), atol=3e-3)


def test_lr_schedule_endpoints():
  """math_test.py:72-154."""
  assert abs(O.learning_rate_decay(0, 2e-3, 2e-5, 1000) - 2e-3) < 1e-12
  assert abs(O.learning_rate_decay(1000, 2e-3, 2e-5, 1000) - 2e-5) < 1e-12
  assert abs(O.learning_rate_decay(0, 2e-3, 2e-5, 1000, 512, 0.01) - 2e-5) < 1e-12


def test_model_shapes_and_param_count():
  """generate_tables.ipynb:232,408 param counts via the layer formula."""
  cfg = O.ModelConfig(num_levels=3, nerf_mlp=O.MLPConfig(net_width=1024, warp_fn='contract'),
                      prop_mlp=O.MLPConfig(net_depth=4, net_width=256, disable_rgb=True, warp_fn='contract'))
  n = sum(fi * fo + fo for m, g in ((cfg.nerf_mlp, 0), (cfg.prop_mlp, 0)) for _, (fi, fo) in O.mlp_param_shapes(m, g))
  assert n == 9007493


def test_model_forward_and_train_step_smoke():
  cfg = O.ModelConfig(num_levels=2, num_prop_samples=16, num_nerf_samples=32, raydist_fn='reciprocal',
                      opaque_background=True,
                      nerf_mlp=O.MLPConfig(net_width=64, warp_fn='contract'),
                      prop_mlp=O.MLPConfig(net_depth=2, net_width=32, disable_rgb=True, warp_fn='contract'))
  import os
  basis = torch.tensor(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geopoly_basis.npz'))['icosahedron_2'].T,
                       dtype=torch.float32)
  params = O.init_params(cfg, seed=0)
  rng = np.random.default_rng(0)
  B = 8
  d = rng.normal(size=(B, 3)).astype(np.float32)
  rays = dict(origins=torch.tensor(rng.normal(size=(B, 3)).astype(np.float32) * 0.3),
              directions=torch.tensor(d), viewdirs=torch.tensor(d / np.linalg.norm(d, axis=-1, keepdims=True)),
              radii=torch.full((B, 1), 1e-3), near=torch.full((B, 1), 0.2), far=torch.full((B, 1), 1e6),
              lossmult=torch.ones(B, 1), static_mask=torch.ones(B, 1), embed_idx=torch.zeros(B, 1, dtype=torch.int32))
  rend, hist = O.model_apply(cfg, params, rays, 0.5, True, basis)
  assert rend[-1]['rgb'].shape == (B, 3) and hist[0]['sdist'].shape == (B, 17) and hist[1]['weights'].shape == (B, 32)
  assert all(torch.isfinite(v).all() for v in rend[-1].values())
  gt = torch.tensor(rng.uniform(size=(B, 3)).astype(np.float32))
  p2, opt, stats, grads = O.train_step(cfg, O.LossConfig(), params, O.init_opt_state(params), 0, rays, gt, 0.5, basis)
  assert np.isfinite(float(stats['loss']))
  assert any(float(g.abs().max()) > 0 for n, g in grads.items() if n.startswith('PropMLP_0'))
  assert any(float(g.abs().max()) > 0 for n, g in grads.items() if n.startswith('NerfMLP_0'))


def test_adam_update_matches_an_independent_implementation():
  """optax.adam is not importable here (parity unpinned for it); its published update rule
  m_hat / (sqrt(v_hat) + eps) with bias correction (train_utils.py:487-512, eps = 1e-6) is the one torch.optim.Adam
  implements, so the oracle's update is checked against that independent implementation on the oracle's own gradients
  (gradient clipping switched off so that the raw gradients are what both optimisers see)."""
  from tests import helpers as H
  torch.manual_seed(0)
  ocfg, _ = H.config_pair(n_prop=16, n_nerf=16, max_rays=8)
  lcfg = O.LossConfig(grad_max_norm=0.0, grad_max_val=0.0)
  basis = torch.tensor(H.basis_np())
  params = O.init_params(ocfg, seed=0, bias_scale=0.1)
  rays, gt = H.make_rays(8, seed=1)
  names = [n for n, _ in O.tree_leaves(params)]
  twin = {n: torch.nn.Parameter(v.detach().clone()) for n, v in O.tree_leaves(params)}
  opt = torch.optim.Adam(list(twin.values()), lr=1.0, betas=(lcfg.adam_beta1, lcfg.adam_beta2), eps=lcfg.adam_eps)
  state = O.init_opt_state(params)
  for step in range(3):
    params, state, stats, raw = O.train_step(ocfg, lcfg, params, state, step, rays, gt, 0.5, basis)
    for g in opt.param_groups:
      g['lr'] = stats['lr']
    for n in names:
      twin[n].grad = raw[n].detach().clone()
    opt.step()
  for n, v in O.tree_leaves(params):
    np.testing.assert_allclose(v.detach().numpy(), twin[n].detach().numpy(), rtol=2e-6, atol=1e-9, err_msg=n)
