"""Runs a few render forwards / training steps of config A (target process for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mipnerf360 as O
from tests import helpers as H
from tests.test_gpu_train import _loss_cfg
from nerf_hugs_b200.engine import Engine

mode = sys.argv[1] if len(sys.argv) > 1 else 'render'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 4096
ocfg, ecfg = H.config_pair(precision='bf16_tc', max_rays=n)
params = O.init_params(ocfg, seed=0)
rays, gt = H.make_rays(n, seed=1)
eng = Engine(ecfg, H.basis_np())
flat = eng.flatten_params(params); eng.params_changed(flat)
dev = flat.device
rays = {k: v.to(dev) for k, v in rays.items()}; gt = gt.to(dev)
jit = torch.rand(2, n, device=dev)
lc = _loss_cfg(O.LossConfig())
grad = torch.empty_like(flat); stats = torch.empty(16, device=dev)
for _ in range(iters):
  if mode == 'render':
    eng.forward(flat, rays, 0.5, None, compute_extras=True, want_history=False)
  else:
    eng.loss_and_grad(flat, rays, gt, 0.5, jit, lc, grad, stats)
torch.cuda.synchronize()
print('done', mode, iters)
