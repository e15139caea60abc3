"""Golden fixtures of the vanilla-NeRF torch twin, produced by RUNNING THE REFERENCE ITSELF.

`/root/reference/nerfacto/models/nerf.py` (+ utils/ray_utils.py, models/custom_functions.py) is imported unmodified and
run on the CPU; `tinycudann` (needed only because models/__init__.py imports nerfacto.py) is an empty stub module.
The reference tree does not exist on the GPU box, so the outputs are committed as tests/golden/nerfacto_nerf.npz.

Per case: the batch, the uniform draws the reference consumed (torch.rand is intercepted so that the CUDA path can be run
on the same draws), every output of Model.forward, the loss / info_dict of Loss.forward, and of loss.backward() the full
gradient of the small tensors plus {L2 norm, projection on a fixed pseudo-random vector} of every tensor.  Initial weights
are NOT stored: the product's parameter containers draw them from the same seed in the same order; a checksum pins that.

Run once in the build container:  python tests/golden/make_golden_nerfacto_nerf.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/nerfacto'
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
  sys.modules.setdefault('tinycudann', types.ModuleType('tinycudann'))
  sys.path.insert(0, REF)
  try:
    import models as ref_models            # noqa: the reference's own package
    from models import nerf as ref_nerf
  finally:
    sys.path.remove(REF)
  return ref_models, ref_nerf


def make_batch(n_rays, seed, appearance=False, mask=False):
  """BASELINE config 1 geometry: pinhole focal 70, 64x64 image, camera at (0, 0, 4) looking at the origin."""
  g = torch.Generator().manual_seed(seed)
  H = W = 64
  focal = 70.
  pix = torch.randint(0, H * W, (n_rays,), generator=g)
  py, px = (pix // W).float(), (pix % W).float()
  dirs = torch.stack([(px + 0.5 - W / 2) / focal, -(py + 0.5 - H / 2) / focal, -torch.ones(n_rays)], -1)
  origin = torch.tensor([0., 0., 4.]).expand(n_rays, 3).contiguous()
  viewdir = dirs / dirs.norm(dim=-1, keepdim=True)
  batch = {
      'coord': torch.stack([px / W, py / H], -1),
      'origin': origin, 'direction': dirs.contiguous(), 'viewdir': viewdir.contiguous(),
      'bg_rgb': torch.rand(n_rays, 3, generator=g),
      'embed_idx': torch.randint(0, 30, (n_rays, 1), generator=g).int(),
      'near': torch.full((n_rays, 1), 2.0), 'far': torch.full((n_rays, 1), 6.0),
      'rgb': torch.rand(n_rays, 3, generator=g),
      'static_mask': (torch.rand(n_rays, 1, generator=g) < 0.8).float(),
  }
  return batch


CASES = {
    # BASELINE config 1: kubric_nerf_base.yml model section
    'cfg1': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=False, eval_embedding='original',
                            opaque_background=True, num_coarse_nerf_samples_per_ray=64, num_fine_nerf_samples_per_ray=64,
                            proposal_initial_sampler='uniform', rgb_loss_type='mse'),
                 n_rays=256, contraction=False, perturb=True, train=True, seed=0),
    # phototourism_nerf_base.yml shape: appearance embedding 48, charbonnier, contraction + reciprocal spacing for coverage
    'photo': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=True, appearance_embedding_dim=48,
                             num_embedding=30, eval_embedding='original', opaque_background=False,
                             num_coarse_nerf_samples_per_ray=32, num_fine_nerf_samples_per_ray=48,
                             proposal_initial_sampler='reciprocal', rgb_loss_type='charb', use_single_jitter=True),
                  n_rays=96, contraction=True, perturb=True, train=True, seed=1),
    # config 1's model at a 4096-ray batch (the size at which the gradient bar of 1e-3 is meaningful for a ReLU network: see
    # DESIGN.md "What bounds a gradient comparison"); compact record: batch regenerated from its seed, no per-sample arrays
    'cfg1_4096': dict(model=dict(net_width=256, max_deg_point=15, use_appearance_embedding=False, eval_embedding='original',
                                 opaque_background=True, num_coarse_nerf_samples_per_ray=64, num_fine_nerf_samples_per_ray=64,
                                 proposal_initial_sampler='uniform', rgb_loss_type='mse', use_single_jitter=True),
                      n_rays=4096, contraction=False, perturb=True, train=True, seed=3, compact=True),
    # evaluation path: deterministic sampling, average embedding, chunked
    'eval': dict(model=dict(net_width=256, max_deg_point=12, use_appearance_embedding=True, appearance_embedding_dim=8,
                            num_embedding=30, eval_embedding='average', opaque_background=True,
                            num_coarse_nerf_samples_per_ray=16, num_fine_nerf_samples_per_ray=24,
                            proposal_initial_sampler='piecewise'),
                 n_rays=80, contraction=True, perturb=False, train=False, seed=2),
}


def projection_vector(numel, tag):
  rng = np.random.default_rng(abs(hash_name(tag)) % (2 ** 32))
  return rng.standard_normal(numel).astype(np.float32)


def hash_name(s):
  h = 2166136261
  for ch in s.encode():
    h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
  return h


def main():
  ref_models, ref_nerf = import_reference()
  out = {}
  real_rand = torch.rand
  for name, case in CASES.items():
    torch.manual_seed(1234 + case['seed'])
    cfg = ref_nerf.ModelConfig(**case['model'])
    model = ref_models.model_dict['nerf'](cfg, 1.0, False, case['contraction'])
    crit = ref_models.criterion_dict['nerf'](model)
    sd = model.state_dict()
    out[f'{name}/weights_checksum'] = np.array([float(sum(v.double().abs().sum() for v in sd.values())),
                                                float(sum(v.numel() for v in sd.values()))])
    batch = make_batch(case['n_rays'], case['seed'])
    compact = case.get('compact', False)
    if compact:
      out[f'{name}/batch_checksum'] = np.array(float(sum(v.double().abs().sum() for v in batch.values())))
    else:
      for k, v in batch.items():
        out[f'{name}/batch/{k}'] = v.numpy()
    draws = []

    def rand_spy(*a, **k):
      r = real_rand(*a, **k)
      draws.append(r.clone())
      return r
    model.train(case['train'])
    torch.rand = rand_spy
    # the fenceposts and weights of every field evaluation (one per field and chunk), to separate sampling parity from
    # field / compositing parity in the tests
    levels = []
    real_d2w = ref_nerf.density_to_weight

    def d2w_spy(densities, euclidean_bins, directions, opaque_background=False):
      res = real_d2w(densities, euclidean_bins, directions, opaque_background)
      levels.append((euclidean_bins.detach().clone(), res[0].detach().clone()))
      return res
    ref_nerf.density_to_weight = d2w_spy
    try:
      if case['train']:
        outputs = model(batch=batch, curr_step=1, perturb=case['perturb'])
      else:
        with torch.no_grad():
          outputs = model(batch=batch, curr_step=1, perturb=case['perturb'], chunk_size=32)
    finally:
      torch.rand = real_rand
      ref_nerf.density_to_weight = real_d2w
    n_chunks = len(levels) // 2
    for f, ft in enumerate(() if compact else ('coarse', 'fine')):
      out[f'{name}/bins/{ft}'] = torch.cat([levels[2 * c + f][0] for c in range(n_chunks)]).numpy()
      out[f'{name}/weights/{ft}'] = torch.cat([levels[2 * c + f][1] for c in range(n_chunks)]).numpy()
    for i, dr in enumerate(draws):
      out[f'{name}/jitter/{i}'] = dr.numpy()
    out[f'{name}/n_jitter'] = np.array(len(draws))
    for k, v in outputs.items():
      out[f'{name}/out/{k}'] = v.detach().numpy()
    if case['train']:
      n = case['n_rays']
      loss, info, _ = crit(outputs=outputs, batch=batch, data_shape=(n // 16, 4, 4), is_finetune=False, extra_infos={})
      loss.backward()
      out[f'{name}/loss'] = np.array(float(loss))
      for k, v in info.items():
        out[f'{name}/info/{k}'] = np.array(float(v))
      for pname, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        g = g.detach().reshape(-1).numpy()
        out[f'{name}/gsum/{pname}'] = np.array([np.linalg.norm(g.astype(np.float64)),
                                                float(g.astype(np.float64) @ projection_vector(g.size, pname))])
        if g.size <= 93 * 256:
          out[f'{name}/grad/{pname}'] = p.grad.detach().numpy() if p.grad is not None else np.zeros(p.shape, np.float32)
  path = os.path.join(OUT, 'nerfacto_nerf.npz')
  np.savez_compressed(path, **out)
  print(path, os.path.getsize(path) // 1024, 'KiB', len(out), 'arrays')


if __name__ == '__main__':
  main()
