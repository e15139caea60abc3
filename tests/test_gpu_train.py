"""Training-step parity: hugs_loss_and_grad + hugs_adam_step vs the oracle's autograd train_step.

The tensor-core path computes in bf16 x bf16 -> fp32; the oracle is run with bf16-rounded Dense operands
(quant='bf16').  Gradients flow through saved bf16 activations and bf16 dZ tiles, so they are compared by
relative L2 error / cosine per tensor; losses (fp32 compositing + loss kernels) are compared tightly.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O
from tests import helpers as H
from tests.test_gpu_model import _report

pytestmark = pytest.mark.gpu


def _loss_cfg(lcfg: O.LossConfig):
  from nerf_hugs_b200 import _lib
  c = _lib.LossCfg()
  c.data_loss_type = 0 if lcfg.data_loss_type == 'charb' else 1
  c.charb_padding, c.data_loss_mult = lcfg.charb_padding, lcfg.data_loss_mult
  c.data_coarse_loss_mult = lcfg.data_coarse_loss_mult
  c.interlevel_loss_mult, c.distortion_loss_mult = lcfg.interlevel_loss_mult, lcfg.distortion_loss_mult
  c.use_static_mask = int(lcfg.transient_type == 'withmask')
  c.withmask_transient_weight = lcfg.withmask_transient_weight
  c.disable_multiscale_loss = int(lcfg.disable_multiscale_loss)
  return c


def _setup(n=64, glo=0, transient=None, num_levels=2, seed=0, n_prop=64, n_nerf=128):
  from nerf_hugs_b200.engine import Engine
  ocfg, ecfg = H.config_pair(num_levels=num_levels, n_prop=n_prop, n_nerf=n_nerf, precision='bf16_tc',
                             max_rays=max(n, 128), glo=glo)
  lcfg = O.LossConfig(transient_type=transient, distortion_loss_mult=0.01, interlevel_loss_mult=1.0)
  basis = H.basis_np()
  params = O.init_params(ocfg, seed=seed, bias_scale=0.1)
  rays, gt = H.make_rays(n, seed=seed + 1)
  g = torch.Generator().manual_seed(11)
  jit = [torch.rand(n, 1, generator=g) for _ in range(num_levels)]
  eng = Engine(ecfg, basis)
  return ocfg, lcfg, basis, params, rays, gt, jit, eng


def _grad_tree(eng, flat_grad):
  return {name: flat_grad[off:off + r * c].cpu() for name, off, r, c, _ in eng.layout}


@pytest.mark.parametrize('glo,transient', [(0, None), (4, 'withmask')])
def test_loss_and_grad_vs_oracle(glo, transient):
  ocfg, lcfg, basis, params, rays, gt, jit, eng = _setup(glo=glo, transient=transient)
  _, _, stats, ref_grads = O.train_step(ocfg, lcfg, params, O.init_opt_state(params), 0, rays, gt, 0.6,
                                        torch.tensor(basis), jitter=jit, quant='bf16')
  flat = eng.flatten_params(params)
  eng.params_changed(flat)
  jt = torch.stack([j[:, 0] for j in jit])
  grad, st = eng.loss_and_grad(flat, rays, gt, 0.6, jt, _loss_cfg(lcfg))
  torch.cuda.synchronize()
  st = st.cpu().numpy()
  ref_losses = {k: float(v) for k, v in stats['losses'].items()}
  rep = {'loss': [float(st[0]), float(stats['loss'])], 'data': [float(st[1]), ref_losses['data']],
         'interlevel': [float(st[2]), ref_losses['interlevel']], 'distortion': [float(st[3]), ref_losses['distortion']]}
  np.testing.assert_allclose(st[1], ref_losses['data'], rtol=5e-3)
  np.testing.assert_allclose(st[2], ref_losses['interlevel'], rtol=5e-2, atol=1e-6)
  np.testing.assert_allclose(st[3], ref_losses['distortion'], rtol=2e-2, atol=1e-7)
  got = _grad_tree(eng, grad)
  bad = []
  for name, _, r, c, _ in eng.layout:
    ref = ref_grads[name].reshape(-1)
    g = got[name]
    assert torch.isfinite(g).all(), name
    rn = float(ref.norm())
    if rn < 1e-12:
      assert float(g.norm()) < 1e-8, name
      continue
    rel = float((g - ref).norm() / rn)
    cos = float((g * ref).sum() / (g.norm() * ref.norm() + 1e-30))
    rep[name] = [rel, cos]
    lim = (0.3, 0.95) if 'GloEmbed' in name else (0.15, 0.99)
    if not (rel < lim[0] and cos > lim[1]):
      bad.append((name, rel, cos))
  _report(f'loss_and_grad_glo{glo}_{transient}', rep)
  eng.close()
  assert not bad, bad


def test_adam_step_vs_oracle():
  """clip_gradients + nan_to_num + optax.adam on identical gradients: params after 2 steps match to 1e-6."""
  from nerf_hugs_b200 import _lib
  ocfg, lcfg, basis, params, rays, gt, jit, eng = _setup(n=16, n_prop=16, n_nerf=32)
  opt = O.init_opt_state(params)
  flat = eng.flatten_params(params)
  mu, nu = torch.zeros_like(flat), torch.zeros_like(flat)
  p_ref = params
  for step in range(2):
    p_ref, opt, stats, raw = O.train_step(ocfg, lcfg, p_ref, opt, step, rays, gt, 0.5, torch.tensor(basis),
                                          jitter=jit, quant=None)
    gflat = eng.flatten_params(_nest(raw))
    if step == 1:
      gflat[7] = float('nan')          # nan_to_num path (train_utils.py:466)
    a = _lib.AdamCfg()
    a.lr, a.beta1, a.beta2, a.eps = stats['lr'], lcfg.adam_beta1, lcfg.adam_beta2, lcfg.adam_eps
    a.grad_max_norm, a.grad_max_val, a.step, a.grad_scale = lcfg.grad_max_norm, lcfg.grad_max_val, step, 1.0
    norms = torch.zeros(9, device=flat.device)
    eng.adam_step(flat, gflat, mu, nu, a, norms)
    torch.cuda.synchronize()
    if step == 0:
      ref_flat = eng.flatten_params(p_ref)
      np.testing.assert_allclose(flat.cpu().numpy(), ref_flat.cpu().numpy(), rtol=2e-5, atol=1e-7)
      # grad norm of NerfMLP_0 (stats['grad_norms'])
      nerf = torch.cat([g.reshape(-1) for n, g in raw.items() if n.startswith('NerfMLP_0')])
      np.testing.assert_allclose(float(norms[0]), float(nerf.norm()), rtol=1e-4)
  assert torch.isfinite(flat).all()
  eng.close()


def _nest(flat_named):
  tree = {}
  for name, v in flat_named.items():
    parts = name.split('/')
    d = tree
    for q in parts[:-1]:
      d = d.setdefault(q, {})
    d[parts[-1]] = v
  return tree


def test_training_reduces_loss():
  """A few optimisation steps on a fixed batch through the public engine calls: loss goes down."""
  from nerf_hugs_b200 import _lib
  ocfg, lcfg, basis, params, rays, gt, jit, eng = _setup(n=128)
  flat = eng.flatten_params(params)
  eng.params_changed(flat)
  mu, nu = torch.zeros_like(flat), torch.zeros_like(flat)
  jt = torch.stack([j[:, 0] for j in jit])
  lc = _loss_cfg(lcfg)
  losses = []
  for step in range(12):
    grad, st = eng.loss_and_grad(flat, rays, gt, 0.5, jt, lc)
    a = _lib.AdamCfg()
    a.lr, a.beta1, a.beta2, a.eps = 2e-3, 0.9, 0.999, 1e-6
    a.grad_max_norm, a.grad_max_val, a.step, a.grad_scale = 0.0, 0.0, step, 1.0
    eng.adam_step(flat, grad, mu, nu, a)
    losses.append(float(st[0]))
  assert all(np.isfinite(losses)), losses
  assert losses[-1] < losses[0], losses
  _report('training_losses', losses)
  eng.close()
