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
    'cfg1_4096': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=False, eval_embedding='original',
                                 opaque_background=True, num_coarse_nerf_samples_per_ray=64, num_fine_nerf_samples_per_ray=64,
                                 proposal_initial_sampler='uniform', rgb_loss_type='mse', use_single_jitter=True),
                      n_rays=4096, contraction=False, perturb=True, train=True, seed=3, compact=True),
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


def make_batch(n_rays, seed):
  """== make_batch of tests/golden/make_golden_nerfacto_nerf.py (BASELINE config 1 geometry)."""
  g = torch.Generator().manual_seed(seed)
  H = W = 64
  focal = 70.
  pix = torch.randint(0, H * W, (n_rays,), generator=g)
  py, px = (pix // W).float(), (pix % W).float()
  dirs = torch.stack([(px + 0.5 - W / 2) / focal, -(py + 0.5 - H / 2) / focal, -torch.ones(n_rays)], -1)
  origin = torch.tensor([0., 0., 4.]).expand(n_rays, 3).contiguous()
  viewdir = dirs / dirs.norm(dim=-1, keepdim=True)
  return {
      'coord': torch.stack([px / W, py / H], -1),
      'origin': origin, 'direction': dirs.contiguous(), 'viewdir': viewdir.contiguous(),
      'bg_rgb': torch.rand(n_rays, 3, generator=g),
      'embed_idx': torch.randint(0, 30, (n_rays, 1), generator=g).int(),
      'near': torch.full((n_rays, 1), 2.0), 'far': torch.full((n_rays, 1), 6.0),
      'rgb': torch.rand(n_rays, 3, generator=g),
      'static_mask': (torch.rand(n_rays, 1, generator=g) < 0.8).float(),
  }


def load_batch(gold, name, device=None):
  if CASES[name].get('compact'):
    batch = make_batch(CASES[name]['n_rays'], CASES[name]['seed'])
    got = float(sum(v.double().abs().sum() for v in batch.values()))
    want = float(gold[f'{name}/batch_checksum'])
    assert abs(got - want) <= 1e-9 * abs(want), 'the seeded batch differs from the one the reference was run on'
    return {k: (v.to(device) if device is not None else v) for k, v in batch.items()}
  batch = {}
  for k in gold.files:
    if k.startswith(f'{name}/batch/'):
      t = torch.from_numpy(gold[k])
      batch[k.split('/')[-1]] = t.to(device) if device is not None else t
  return batch


# ------------------------------------------------------------------------------------------------ nerfacto (hash grid)
GOLDEN_HASH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nerfacto_hash.npz')
PROP_ARGS = [
    {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 12, 'features_per_level': 2, 'num_levels': 5, 'max_res': 64},
    {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 13, 'features_per_level': 2, 'num_levels': 7, 'max_res': 128},
]
FIELD = dict(hidden_dim=256, geo_feat_dim=64, hidden_dim_color=256, base_res=16, max_res=512, log2_hashmap_size=15,
             features_per_level=2, enable_tcnn_mlp=False, num_levels=16)
GRID_GAIN = 3000.0

# must equal CASES of tests/golden/make_golden_nerfacto.py
HASH_CASES = {
    'withmask': dict(model=dict(**FIELD, transient_type='withmask', use_appearance_embedding=True, use_transient_embedding=False,
                                appearance_embedding_dim=48, num_embedding=30, eval_embedding='original', opaque_background=True,
                                num_nerf_samples_per_ray=16, num_proposal_samples_per_ray=(32, 24), num_proposal_iterations=2,
                                proposal_net_args_list=PROP_ARGS, proposal_initial_sampler='uniform',
                                proposal_histogram_padding=0.005, proposal_weights_anneal_max_num_iters=10000,
                                rgb_loss_type='charb', distortion_loss_mult=0.001),
                     n_rays=128, bound=2.0, contraction=False, perturb=True, train=True, step=40, seed=5),
    'contract': dict(model=dict(**FIELD, transient_type=None, use_appearance_embedding=False, opaque_background=False,
                                density_activation='softplus', num_nerf_samples_per_ray=24,
                                num_proposal_samples_per_ray=(40, 28), num_proposal_iterations=2, use_same_proposal_network=True,
                                proposal_net_args_list=PROP_ARGS[:1], proposal_initial_sampler='piecewise',
                                rgb_loss_type='mse', use_single_jitter=False),
                     n_rays=96, bound=2.0, contraction=True, perturb=True, train=True, step=2000, seed=6),
    'eval': dict(model=dict(**FIELD, transient_type='withmask', use_appearance_embedding=True, appearance_embedding_dim=8,
                            num_embedding=30, eval_embedding='average', opaque_background=True, num_nerf_samples_per_ray=16,
                            num_proposal_samples_per_ray=(32,), num_proposal_iterations=1, proposal_net_args_list=PROP_ARGS[1:],
                            proposal_initial_sampler='uniform'),
                 n_rays=80, bound=2.0, contraction=False, perturb=False, train=False, step=500, seed=7),
}


def build_hash(name, device=None, precision=None):
  from nerf_hugs_b200.nerfacto.models import criterion_dict, model_config_dict, model_dict
  case = HASH_CASES[name]
  torch.manual_seed(4321 + case['seed'])
  cfg = model_config_dict['nerfacto'](**case['model'])
  model = model_dict['nerfacto'](cfg, case['bound'], False, case['contraction'])
  if precision is not None:
    model.precision = precision
  crit = criterion_dict['nerfacto'](model)
  with torch.no_grad():
    for pname, p in model.named_parameters():
      if pname.endswith('mlp_base.0.params'):
        p.mul_(GRID_GAIN)
  if device is not None:
    model = model.to(device)
  return case, model, crit


def load_hash_batch(gold, name, device=None):
  batch = {}
  for k in gold.files:
    if k.startswith(f'{name}/batch/'):
      t = torch.from_numpy(gold[k])
      batch[k.split('/')[-1]] = t.to(device) if device is not None else t
  return batch
