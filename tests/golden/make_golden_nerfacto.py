"""Golden fixtures of the nerfacto torch twin, produced by RUNNING THE REFERENCE'S OWN nerfacto.py.

`/root/reference/nerfacto/models/nerfacto.py` (+ utils/ray_utils.py, utils/loss_utils.py, models/custom_functions.py) is
imported unmodified and run on the CPU in float32 with `enable_tcnn_mlp: False` (as every shipped yml).  Its only
dependency that is not under the reference tree, tiny-cuda-nn, is replaced by oracle/hashgrid.py - a restatement of tcnn's
published HashGrid / SphericalHarmonics encodings (PARITY UNPINNED against tcnn itself, see that file).  Everything else
- the proposal loop, annealing, sampling, MLPs, compositing, the three losses, autograd - is the reference's code.

The reference tree does not exist on the GPU box: outputs are committed as tests/golden/nerfacto_hash.npz.
Run once in the build container:  python tests/golden/make_golden_nerfacto.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hashgrid as tcnn_shim           # noqa: E402
from tests.golden.make_golden_nerfacto_nerf import hash_name, make_batch, projection_vector   # noqa: E402

REF = '/root/reference/nerfacto'
OUT = os.path.dirname(os.path.abspath(__file__))

PROP_ARGS = [
    {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 12, 'features_per_level': 2, 'num_levels': 5, 'max_res': 64},
    {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 13, 'features_per_level': 2, 'num_levels': 7, 'max_res': 128},
]
FIELD = dict(hidden_dim=256, geo_feat_dim=64, hidden_dim_color=256, base_res=16, max_res=512, log2_hashmap_size=15,
             features_per_level=2, enable_tcnn_mlp=False, num_levels=16)
GRID_GAIN = 3000.0       # tcnn initialises the tables in [-1e-4, 1e-4]: scaled up so that the encodings matter in the test

CASES = {
    # phototourism_nerfacto_withmask.yml shape (BASELINE config 4), small tables
    'withmask': dict(model=dict(**FIELD, transient_type='withmask', use_appearance_embedding=True, use_transient_embedding=False,
                                appearance_embedding_dim=48, num_embedding=30, eval_embedding='original', opaque_background=True,
                                num_nerf_samples_per_ray=16, num_proposal_samples_per_ray=(32, 24), num_proposal_iterations=2,
                                proposal_net_args_list=PROP_ARGS, proposal_initial_sampler='uniform',
                                proposal_histogram_padding=0.005, proposal_weights_anneal_max_num_iters=10000,
                                rgb_loss_type='charb', distortion_loss_mult=0.001),
                     n_rays=128, bound=2.0, contraction=False, perturb=True, train=True, step=40, seed=5),
    # distractor_nerfacto_withmask.yml shape: scene contraction, piecewise sampler, one shared proposal network
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


def import_reference():
  sys.modules['tinycudann'] = tcnn_shim
  sys.path.insert(0, REF)
  try:
    import models as ref_models            # noqa
    from models import nerfacto as ref_nerfacto
  finally:
    sys.path.remove(REF)
  return ref_models, ref_nerfacto


def scale_grids(model):
  with torch.no_grad():
    for name, p in model.named_parameters():
      if name.endswith('mlp_base.0.params'):
        p.mul_(GRID_GAIN)


def main():
  ref_models, ref_nerfacto = import_reference()
  out = {}
  real_rand = torch.rand
  for name, case in CASES.items():
    torch.manual_seed(4321 + case['seed'])
    cfg = ref_nerfacto.ModelConfig(**case['model'])
    model = ref_models.model_dict['nerfacto'](cfg, case['bound'], False, case['contraction'])
    crit = ref_models.criterion_dict['nerfacto'](model)
    scale_grids(model)
    sd = model.state_dict()
    out[f'{name}/weights_checksum'] = np.array([float(sum(v.double().abs().sum() for v in sd.values())),
                                                float(sum(v.numel() for v in sd.values()))])
    out[f'{name}/state_keys'] = np.array(sorted(sd.keys()))
    batch = make_batch(case['n_rays'], case['seed'])
    for k, v in batch.items():
      out[f'{name}/batch/{k}'] = v.numpy()
    draws = []

    def rand_spy(*a, **k):
      r = real_rand(*a, **k)
      draws.append(r.clone())
      return r
    model.train(case['train'])
    torch.rand = rand_spy
    try:
      if case['train']:
        outputs = model(batch=batch, curr_step=case['step'], perturb=case['perturb'])
      else:
        with torch.no_grad():
          outputs = model(batch=batch, curr_step=case['step'], perturb=case['perturb'], chunk_size=32)
    finally:
      torch.rand = real_rand
    for i, dr in enumerate(draws):
      out[f'{name}/jitter/{i}'] = dr.numpy()
    out[f'{name}/n_jitter'] = np.array(len(draws))
    for k, v in outputs.items():
      if isinstance(v, list):
        for i, t in enumerate(v):
          out[f'{name}/out/{k}/{i}'] = t.detach().numpy()
      else:
        out[f'{name}/out/{k}'] = v.detach().numpy()
    if case['train']:
      n = case['n_rays']
      loss, info, _ = crit(outputs=outputs, batch=batch, data_shape=(n // 16, 4, 4), is_finetune=False,
                           extra_infos={'curr_step': case['step']})
      loss.backward()
      out[f'{name}/loss'] = np.array(float(loss.detach()))
      for k, v in info.items():
        out[f'{name}/info/{k}'] = np.array(float(v))
      for pname, p in model.named_parameters():
        if p.numel() == 0:
          continue
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        g = g.detach().reshape(-1).numpy()
        out[f'{name}/gsum/{pname}'] = np.array([np.linalg.norm(g.astype(np.float64)),
                                                float(g.astype(np.float64) @ projection_vector(g.size, pname))])
        if g.size <= 64 * 256 and not pname.endswith('.params'):
          out[f'{name}/grad/{pname}'] = g.reshape(p.shape)
  path = os.path.join(OUT, 'nerfacto_hash.npz')
  np.savez_compressed(path, **out)
  print(path, os.path.getsize(path) // 1024, 'KiB', len(out), 'arrays')


if __name__ == '__main__':
  main()
