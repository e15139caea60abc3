"""Measurements of the torch twins (JSON lines on stdout; one GPU, or N GPUs of one box under
`python -m torch.distributed.run --nproc-per-node N scripts/nerfacto_bench.py ...`: rays shard over the ranks, gradients are
all-reduced, times are the maximum over ranks, rank 0 prints):

  config 1  vanilla NeRF (nerfacto/models/nerf.py, kubric_nerf_base.yml: 64 + 64 samples, pos_enc degree 15, MSE) - the
            reference's CPU plumbing case, here at its yml batch of 4096 rays on the GPU (and at 256 rays)
  config 4  nerfacto hash grid (phototourism_nerfacto_withmask.yml: 2^21 x 16 x 2 table, 256-wide MLPs, proposal
            networks 512 / 256 samples, 128 field samples, appearance embedding 48, HuGS static masks, charbonnier,
            batch 16384 = 64 patches of 16 x 16)

One step = model(batch) -> criterion -> loss.backward() -> torch.optim.Adam.step(), the loop body of nerfacto/train.py:186-208.
Usage: python scripts/nerfacto_bench.py [nerf] [nerfacto] [--steps K] [--rays N]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from nerf_hugs_b200 import _lib
from nerf_hugs_b200.nerfacto.models import criterion_dict, model_config_dict, model_dict
from nerf_hugs_b200.nerfacto.parallel import allreduce_gradients

RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))


def device():
  torch.cuda.set_device(LOCAL)
  dev = torch.device('cuda', LOCAL)
  if WORLD > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  return dev


def emit(d):
  if RANK == 0:
    print(json.dumps(d), flush=True)


def synthetic_batch(n_rays, dev, seed, bound):
  """Phototourism-shape synthetic rays: cameras on a sphere of radius 1.5 * bound looking at the origin, 16 x 16 patches."""
  g = torch.Generator().manual_seed(seed)
  n_patch = max(1, n_rays // 256)
  pos = torch.randn(n_patch, 3, generator=g)
  pos = pos / pos.norm(dim=-1, keepdim=True) * (1.5 * bound)
  fwd = -pos / pos.norm(dim=-1, keepdim=True)
  up = torch.tensor([0., 0., 1.]).expand_as(fwd)
  right = torch.linalg.cross(fwd, up); right = right / (right.norm(dim=-1, keepdim=True) + 1e-9)
  up = torch.linalg.cross(right, fwd)
  px = (torch.rand(n_patch, 1, 1, generator=g) * 0.8 - 0.4) + (torch.arange(16).float()[None, None, :] / 800.)
  py = (torch.rand(n_patch, 1, 1, generator=g) * 0.8 - 0.4) + (torch.arange(16).float()[None, :, None] / 800.)
  d = fwd[:, None, None, :] + px[..., None] * right[:, None, None, :] + py[..., None] * up[:, None, None, :]
  d = d.expand(n_patch, 16, 16, 3).reshape(-1, 3)[:n_rays]
  o = pos[:, None, None, :].expand(n_patch, 16, 16, 3).reshape(-1, 3)[:n_rays]
  n = o.shape[0]
  batch = {
      'coord': torch.rand(n, 2, generator=g), 'origin': o.contiguous(), 'direction': d.contiguous(),
      'viewdir': (d / d.norm(dim=-1, keepdim=True)).contiguous(), 'bg_rgb': torch.rand(n, 3, generator=g),
      'embed_idx': torch.randint(0, 800, (n_patch, 1), generator=g).repeat_interleave(256, 0)[:n].int(),
      'near': torch.full((n, 1), 0.5 * bound), 'far': torch.full((n, 1), 2.5 * bound), 'rgb': torch.rand(n, 3, generator=g),
      'static_mask': (torch.rand(n_patch, 1, generator=g) < 0.8).float().repeat_interleave(256, 0)[:n],
  }
  return {k: v.to(dev) for k, v in batch.items()}


def run(kind, n_rays, steps, warmup=5):
  dev = device()
  torch.manual_seed(0)
  if kind == 'nerf':
    cfg = model_config_dict['nerf'](net_width=256, max_deg_point=15, use_appearance_embedding=False, eval_embedding='original',
                                    opaque_background=True, num_coarse_nerf_samples_per_ray=64, num_fine_nerf_samples_per_ray=64,
                                    proposal_initial_sampler='uniform', rgb_loss_type='mse')
    bound, lr, eps = 1.0, 1e-3, 1e-8
    model = model_dict['nerf'](cfg, bound, False, False).to(dev)
    crit = criterion_dict['nerf'](model)
    samples = {'coarse field': 64, 'fine field': 128}
    label = 'BASELINE config 1: vanilla NeRF (nerfacto/models/nerf.py, kubric_nerf_base.yml: 64 + 64 samples, 256-wide MLPs on the tcgen05 chain kernel)'
  else:
    cfg = model_config_dict['nerfacto'](
        hidden_dim=256, geo_feat_dim=64, hidden_dim_color=256, base_res=16, max_res=8192, log2_hashmap_size=21,
        features_per_level=2, enable_tcnn_mlp=False, transient_type='withmask', use_appearance_embedding=True,
        use_transient_embedding=False, appearance_embedding_dim=48, eval_embedding='original', opaque_background=True,
        num_nerf_samples_per_ray=128, num_proposal_samples_per_ray=(512, 256), num_proposal_iterations=2,
        proposal_net_args_list=[
            {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 17, 'features_per_level': 2, 'num_levels': 5, 'max_res': 512},
            {'base_res': 16, 'hidden_dim': 64, 'log2_hashmap_size': 17, 'features_per_level': 2, 'num_levels': 7, 'max_res': 2048}],
        proposal_initial_sampler='uniform', proposal_histogram_padding=0.005, proposal_weights_anneal_max_num_iters=10000,
        rgb_loss_type='charb', distortion_loss_mult=0.001)
    bound, lr, eps = 2.0, 1e-2, 1e-15
    model = model_dict['nerfacto'](cfg, bound, True, False).to(dev)
    crit = criterion_dict['nerfacto'](model)
    samples = {'proposal 0': 512, 'proposal 1': 256, 'field': 128}
    label = ('BASELINE config 4: nerfacto hash grid (phototourism_nerfacto_withmask.yml: 2^21 x 16 x 2 table, 256-wide field '
             'MLPs, 512 / 256 proposal + 128 field samples, appearance embedding 48, HuGS static masks, charbonnier)')
  per = n_rays // WORLD
  batch = synthetic_batch(per, dev, 11 + RANK, bound)
  opt = torch.optim.Adam([{'params': [p for p in v if p.numel() > 0], 'lr': lr} for v in model.get_params_dict().values()],
                         betas=(0.9, 0.999), eps=eps, fused=True)
  model.train()
  data_shape = (per // 256, 16, 16)
  ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(steps + warmup)] for k in ('s', 'f', 'l', 'b', 'r', 'o')}
  losses = []

  def step(i):
    ev['s'][i].record()
    opt.zero_grad(set_to_none=True)
    outputs = model(batch=batch, curr_step=5000 + i, perturb=True)        # past the proposal warm-up: update every 5th step
    ev['f'][i].record()
    loss, info, _ = crit(outputs=outputs, batch=batch, data_shape=data_shape, is_finetune=False, extra_infos={})
    ev['l'][i].record()
    loss.backward()
    ev['b'][i].record()
    allreduce_gradients(model)
    ev['r'][i].record()
    opt.step()
    ev['o'][i].record()
    return loss

  for i in range(warmup):
    step(i)
  if WORLD > 1:
    dist.barrier()
  torch.cuda.synchronize()
  l0 = _lib.lib.hugs_launch_count()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  for i in range(warmup, warmup + steps):
    loss = step(i)
  lv = float(loss.detach())       # the loss read-back of train.py:215
  t1.record()
  if WORLD > 1:
    dist.barrier()
  torch.cuda.synchronize()
  launches = (_lib.lib.hugs_launch_count() - l0) / steps
  ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev)
  if WORLD > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms)
  rng = range(warmup, warmup + steps)
  phase = lambda a, b: float(np.mean([ev[a][i].elapsed_time(ev[b][i]) for i in rng]))
  n_params = sum(p.numel() for p in model.parameters())
  emit({'config': label, 'metric': 'training rays/s', 'value': n_rays / ms * 1e3, 'ms_per_step': ms, 'rays_per_gpu': per,
        'n_gpus': WORLD, 'steps': steps, 'loss': lv, 'samples_per_ray': samples, 'parameters': n_params,
        'phase_ms': {'forward': phase('s', 'f'), 'loss': phase('f', 'l'), 'backward': phase('l', 'b'),
                     'grad_allreduce': phase('b', 'r'), 'adam(torch fused)': phase('r', 'o')},
        'repo_kernel_launches_per_step': launches, 'optimizer': 'torch.optim.Adam(fused=True), as nerfacto/train.py:153'})


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('what', nargs='*', default=['nerf', 'nerfacto'])
  ap.add_argument('--steps', type=int, default=30)
  ap.add_argument('--rays', type=int, default=0)
  ap.add_argument('--warmup', type=int, default=5)
  a = ap.parse_args()
  if 'nerf' in a.what:
    run('nerf', a.rays or 4096 * WORLD, a.steps, a.warmup)
  if 'nerfacto' in a.what:
    run('nerfacto', a.rays or 16384 * WORLD, a.steps, a.warmup)
  if WORLD > 1:
    dist.barrier()
    dist.destroy_process_group()
