"""NerfMLP.net_width = 512 / 1024 (every non-debug gin of the reference, MipNeRF360/configs/360.gin:14-16) on the
layer-at-a-time tensor-core path (csrc/dense_tc.cu, csrc/layered.cu): forward and training parity against the oracle
that rounds where the kernels round (bf16 operands, bf16 saved activations / dZ), and the shipped gins end to end."""
import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O
from tests import helpers as H
from tests.test_gpu_model import _report, _relerr
from tests.test_gpu_train import _loss_cfg, _grad_tree

pytestmark = pytest.mark.gpu


def _pair(width, n, glo=0, levels=2, n_nerf=128, precision='bf16_tc'):
  from nerf_hugs_b200.engine import Engine
  ocfg, ecfg = H.config_pair(num_levels=levels, n_nerf=n_nerf, nerf_width=width, precision=precision, max_rays=max(n, 128),
                             glo=glo)
  params = O.init_params(ocfg, seed=0, bias_scale=0.1)
  rays, gt = H.make_rays(n, seed=1)
  eng = Engine(ecfg, H.basis_np())
  flat = eng.flatten_params(params)
  eng.params_changed(flat)
  return ocfg, params, rays, gt, eng, flat


@pytest.mark.parametrize('width,n', [(512, 70), (1024, 37), (1024, 260)])
def test_forward_wide_nerf_mlp_vs_bf16_oracle(width, n):
  """Render parity incl. ragged tile counts (n * 128 samples is not a multiple of the 256-row GEMM tile for odd n)."""
  ocfg, params, rays, gt, eng, flat = _pair(width, n)
  with torch.no_grad():
    rend, hist = O.model_apply(ocfg, params, rays, 0.6, True, torch.tensor(H.basis_np()), quant='bf16')
  res, eh = eng.forward(flat, rays, 0.6, None, compute_extras=True)
  torch.cuda.synchronize()
  stats = {k: _relerr(res[-1][k].cpu(), rend[-1][k]) for k in ('rgb', 'acc', 'distance_mean', 'distance_median')}
  _report(f'forward_layered_w{width}_n{n}', stats)
  assert stats['rgb'] < 5e-3 and stats['acc'] < 5e-3 and stats['distance_mean'] < 1e-2, stats
  eng.close()


def test_forward_wide_one_level_per_sample_outputs():
  """One level (identical sample positions): per-sample density / colour of the 1024-wide NerfMLP."""
  ocfg, params, rays, gt, eng, flat = _pair(1024, 48, levels=1)
  with torch.no_grad():
    rend, hist = O.model_apply(ocfg, params, rays, 0.6, True, torch.tensor(H.basis_np()), quant='bf16')
  res, eh = eng.forward(flat, rays, 0.6, None, compute_extras=True)
  torch.cuda.synchronize()
  d_ref = hist[-1]['density']
  assert float((eh[-1]['density'].cpu() - d_ref).abs().max()) < 0.02 * max(1.0, float(d_ref.max()))
  assert float((eh[-1]['rgb'].cpu() - hist[-1]['rgb']).abs().max()) < 0.02
  eng.close()


@pytest.mark.parametrize('width,glo', [(512, 0), (768, 0), (1024, 4)])   # 768: three N tiles, the per-tile atomics path of the column sums
def test_training_wide_nerf_mlp(width, glo):
  """loss + gradients of the wide NerfMLP (and of the 256-wide PropMLP on the chain kernel next to it) against the
  bf16-training oracle; the same tolerances as the chain path's two-level test."""
  n = 64
  ocfg, params, rays, gt, eng, flat = _pair(width, n, glo=glo)
  lcfg = O.LossConfig(distortion_loss_mult=0.01, interlevel_loss_mult=1.0)
  g = torch.Generator().manual_seed(11)
  jit = [torch.rand(n, 1, generator=g) for _ in range(2)]
  _, _, stats, ref = O.train_step(ocfg, lcfg, params, O.init_opt_state(params), 0, rays, gt, 0.6, torch.tensor(H.basis_np()),
                                  jitter=jit, quant='bf16_train')
  grad, st = eng.loss_and_grad(flat, rays, gt, 0.6, torch.stack([j[:, 0] for j in jit]), _loss_cfg(lcfg))
  torch.cuda.synchronize()
  st = st.cpu().numpy()
  np.testing.assert_allclose(st[1], float(stats['losses']['data']), rtol=5e-3)
  np.testing.assert_allclose(st[2], float(stats['losses']['interlevel']), rtol=5e-2, atol=1e-6)
  got = _grad_tree(eng, grad)
  rep, bad = {}, []
  fg, fr = [], []
  for name, _, r, c, _ in eng.layout:
    refv = ref[name].reshape(-1)
    gv = got[name]
    assert torch.isfinite(gv).all(), name
    if float(refv.norm()) < 1e-12:
      continue
    rel = float((gv - refv).norm() / refv.norm())
    cos = float((gv * refv).sum() / (gv.norm() * refv.norm() + 1e-30))
    rep[name] = [rel, cos]
    fg.append(gv); fr.append(refv)
    lim = (0.3, 0.95) if 'GloEmbed' in name else (0.2, 0.98)
    if not (rel < lim[0] and cos > lim[1]):
      bad.append((name, rel, cos))
  fg, fr = torch.cat(fg), torch.cat(fr)
  rep['flat_rel'] = float((fg - fr).norm() / fr.norm())
  _report(f'layered_grads_w{width}_glo{glo}', rep)
  assert not bad, bad
  assert rep['flat_rel'] < 0.05, rep['flat_rel']
  eng.close()


def test_training_wide_one_cta_weight_gradients(monkeypatch):
  # HUGS_WGRAD_PAIRS=0: the one-CTA weight-gradient kernel with its own bias column sums (no column sums in the dgrad
  # epilogue) stays covered; the CTA-pair kernel is the default
  monkeypatch.setenv('HUGS_WGRAD_PAIRS', '0')
  test_training_wide_nerf_mlp(1024, 0)


GIN_360 = """
Config.dataset_loader = 'llff'
Config.near = 0.2
Config.far = 1e6
Config.factor = 4

Model.raydist_fn = @jnp.reciprocal
Model.opaque_background = True

PropMLP.warp_fn = @coord.contract
PropMLP.net_depth = 4
PropMLP.net_width = 256
PropMLP.disable_rgb = True

NerfMLP.warp_fn = @coord.contract
NerfMLP.net_depth = 8
NerfMLP.net_width = 1024
"""

GIN_PHOTOTOURISM_1024_WITHMASK = """
Config.dataset_loader = 'phototourism'
Config.near = 1
Config.far = 2
Config.factor = 2
Config.patch_size = 16
Config.transient_type = 'withmask'
Config.finetune_enable = True
Config.distortion_loss_mult = 0.001

Model.num_glo_features = 48
Model.opaque_background = True

PropMLP.net_depth = 4
PropMLP.net_width = 256
PropMLP.disable_rgb = True

NerfMLP.net_depth = 8
NerfMLP.net_width = 1024
"""


@pytest.mark.parametrize('name,gin', [('360', GIN_360), ('phototourism_1024_withmask', GIN_PHOTOTOURISM_1024_WITHMASK)])
def test_shipped_gins_train_through_train_pstep(tmp_path, name, gin):
  """The bindings of MipNeRF360/configs/360.gin and phototourism_1024_withmask.gin, unmodified (3 levels 64 / 64 / 32,
  NerfMLP 8 x 1024): setup_model -> train_pstep reduces the loss, render_image runs."""
  from nerf_hugs_b200.internal import configs, models, train_utils, utils
  f = tmp_path / f'{name}.gin'
  f.write_text(gin)
  config = configs.load_config([str(f)], ['Config.batch_size = 512', 'Config.render_chunk_size = 512'], save_config=False)
  assert config.bindings.nerf_mlp.net_width == 1024
  model, state, render_eval_pfn, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=512)
  near, far = (0.2, 1e6) if name == '360' else (1.0, 2.0)
  rays, gt = H.make_rays(512, seed=2, near=near, far=far, glo=True, n_embed=32)
  batch = utils.Batch(rays=utils.Rays(**rays), rgb=gt)
  losses = []
  for step in range(6):
    state, stats, _ = train_pstep(7, state, batch, (step + 1) / 100, None)
    losses.append(stats['loss'])
  assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
  H_, W_ = 6, 20
  r2, _ = H.make_rays(H_ * W_, seed=3, near=near, far=far, glo=True, n_embed=32)
  img = utils.Rays(**{k: v.reshape(H_, W_, -1) for k, v in r2.items()})
  rendering = models.render_image(lambda rng, rr: render_eval_pfn(state.params, 0.5, None, rr), img, None, config,
                                  verbose=False)
  assert rendering['rgb'].shape == (H_, W_, 3) and torch.isfinite(rendering['rgb']).all()
  model.engine.close()


# ------------------------------------------------------------------------------------------------------------------
# HUGS_PRECISION_TC_SPLIT on the layer-at-a-time path: the SAME dense_tc_kernel / wgrad_kernel with hi + lo operands
# (A_hi W_hi + A_lo W_hi + A_hi W_lo as K-concatenated segments, fp32 epilogues, hi + lo outputs)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('width,n,levels', [(512, 70, 2), (1024, 37, 2), (1024, 48, 1)])
def test_forward_wide_split_vs_fp32_oracle(width, n, levels):
  """North-star tolerance (1e-4 of the fp32 oracle) for the shipped NerfMLP.net_width = 1024 through the tcgen05 GEMMs."""
  ocfg, params, rays, gt, eng, flat = _pair(width, n, levels=levels, precision='tc_split')
  with torch.no_grad():
    rend, hist = O.model_apply(ocfg, params, rays, 0.6, True, torch.tensor(H.basis_np()))
  res, eh = eng.forward(flat, rays, 0.6, None, compute_extras=True)
  torch.cuda.synchronize()
  stats = {k: _relerr(res[-1][k].cpu(), rend[-1][k]) for k in ('rgb', 'acc', 'distance_mean', 'distance_median')}
  _report(f'forward_layered_split_w{width}_n{n}_L{levels}', stats)
  # distances: rays that end on far samples carry the float32 conditioning of the contracted covariance
  assert stats['rgb'] < 1e-4 and stats['acc'] < 1e-4 and stats['distance_mean'] < 3e-4, stats
  if levels == 1:
    np.testing.assert_allclose(eh[-1]['density'].cpu().numpy(), hist[-1]['density'].numpy(), rtol=2e-3, atol=2e-4)
  eng.close()


@pytest.mark.parametrize('width,glo', [(512, 0), (1024, 4)])
def test_training_wide_split_vs_float64_oracle(width, glo):
  """Gradients of the wide NerfMLP in the split mode against the FLOAT64 oracle: the bound of tests/test_gpu_split.py
  (1e-3 + 2 x dist(float32 oracle, float64 oracle) + 0.025 / sqrt(n_rays): ReLU gate flips, IPE conditioning)."""
  from tests.test_gpu_split import _oracle_grads, _cond_compare, _run_engine
  n = 64
  from nerf_hugs_b200.engine import Engine
  ocfg, ecfg = H.config_pair(num_levels=2, n_nerf=128, nerf_width=width, precision='tc_split', max_rays=128, glo=glo)
  lcfg = O.LossConfig(distortion_loss_mult=0.01, interlevel_loss_mult=1.0)
  params = O.init_params(ocfg, seed=0, bias_scale=0.1)
  rays, gt = H.make_rays(n, seed=1)
  g = torch.Generator().manual_seed(11)
  jit = [torch.rand(n, 1, generator=g) for _ in range(2)]
  eng = Engine(ecfg, H.basis_np())
  ref64, stats = _oracle_grads(ocfg, lcfg, params, rays, gt, jit, torch.float64)
  grad, st = _run_engine(eng, params, rays, gt, jit, lcfg)
  rep = _cond_compare(eng, grad, st, ocfg, lcfg, params, rays, gt, jit, stats, ref64, f'split_grads_layered_w{width}_glo{glo}')
  assert rep['flat_cosine'] > 0.9999, rep
  eng.close()
