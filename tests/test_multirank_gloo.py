"""N > 1 host path on CPU: two gloo ranks exercise the gradient/stats averaging used by train_pstep
(all_reduce(SUM) + grad_scale = 1/world == jax.lax.pmean, train_utils.py:457-459) and the ray sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from nerf_hugs_b200.internal import train_utils, utils
  g = torch.Generator().manual_seed(0)
  batch = torch.rand(16, 3, generator=g)                        # the same global batch on every rank
  mine = utils.rank_slice(batch, rank, world)                   # this rank's rays
  grad = mine.sum(0).repeat(4)                                  # a fake per-rank gradient
  stats = torch.tensor([float(rank + 1), 2.0])
  w = train_utils.allreduce_sum_([grad, stats])
  assert w == world
  names = ['NerfMLP_0/Dense_0/kernel', 'NerfMLP_0/Dense_0/bias']
  st = train_utils._LazyStats(torch.cat([stats, torch.zeros(14 + 9 + 5 * len(names))]), None, 2, w, 1e-3, names)   # host slot: 16 stats + 9 norms + 5 per tensor
  assert 'weight_l2s' in st and set(st['grad_norms']) == {'NerfMLP_0', 'NerfMLP_0/Dense_0', 'NerfMLP_0/Dense_0/kernel', 'NerfMLP_0/Dense_0/bias'}
  np.save(os.path.join(out_dir, f'r{rank}.npy'), np.concatenate([(grad / w).numpy(), [st['loss']]]))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_gradient_mean(tmp_path):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
  r0, r1 = np.load(tmp_path / 'r0.npy'), np.load(tmp_path / 'r1.npy')
  np.testing.assert_array_equal(r0, r1)                         # replicas stay identical
  g = torch.Generator().manual_seed(0)
  batch = torch.rand(16, 3, generator=g)
  expect = (batch[:8].sum(0) + batch[8:].sum(0)).repeat(4) / 2  # pmean of the per-rank gradients
  np.testing.assert_allclose(r0[:-1], expect.numpy(), rtol=1e-6)
  assert abs(r0[-1] - 1.5) < 1e-6                               # pmean of the stats
