"""Shared by the vanilla-NeRF (torch twin) tests: the cases of tests/golden/make_golden_nerfacto_nerf.py."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nerfacto_nerf.npz')

# must equal CASES of tests/golden/make_golden_nerfacto_nerf.py
CASES = {
    'cfg1': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=False, eval_embedding='original',
                            opaque_background=True, num_coarse_nerf_samples_per_ray=64, num_fine_nerf_samples_per_ray=64,
                            proposal_initial_sampler='uniform', rgb_loss_type='mse'),
                 n_rays=256, contraction=False, perturb=True, train=True, seed=0),
    'photo': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=True, appearance_embedding_dim=48,
                             num_embedding=30, eval_embedding='original', opaque_background=False,
                             num_coarse_nerf_samples_per_ray=32, num_fine_nerf_samples_per_ray=48,
                             proposal_initial_sampler='reciprocal', rgb_loss_type='charb', use_single_jitter=True),
                  n_rays=96, contraction=True, perturb=True, train=True, seed=1),
    'eval': dict(model=dict(net_width=256, max_deg_point=12, use_appearance_embedding=True, appearance_embedding_dim=8,
                            num_embedding=30, eval_embedding='average', opaque_background=True,
                            num_coarse_nerf_samples_per_ray=16, num_fine_nerf_samples_per_ray=24,
                            proposal_initial_sampler='piecewise'),
                 n_rays=80, contraction=True, perturb=False, train=False, seed=2),
}


def hash_name(s):
  h = 2166136261
  for ch in s.encode():
    h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
  return h


def projection_vector(numel, tag):
  rng = np.random.default_rng(abs(hash_name(tag)) % (2 ** 32))
  return rng.standard_normal(numel).astype(np.float32)


def build(name, precision=None, device=None):
  """The product's Model / Loss for a golden case, initialised from the seed the reference used."""
  from nerf_hugs_b200.nerfacto.models import criterion_dict, model_config_dict, model_dict
  case = CASES[name]
  torch.manual_seed(1234 + case['seed'])
  cfg = model_config_dict['nerf'](**case['model'])
  model = model_dict['nerf'](cfg, 1.0, False, case['contraction'])
  if precision is not None:
    model.precision = precision
  crit = criterion_dict['nerf'](model)
  if device is not None:
    model = model.to(device)
  return case, model, crit


def load_batch(gold, name, device=None):
  batch = {}
  for k in gold.files:
    if k.startswith(f'{name}/batch/'):
      t = torch.from_numpy(gold[k])
      batch[k.split('/')[-1]] = t.to(device) if device is not None else t
  return batch
