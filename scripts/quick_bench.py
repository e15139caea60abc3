"""Quick phase timing on one GPU (development aid; bench.py is the contract benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mipnerf360 as O
from tests import helpers as H
from tests.test_gpu_train import _loss_cfg
from nerf_hugs_b200.engine import Engine
from nerf_hugs_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ocfg, ecfg = H.config_pair(precision='bf16_tc', max_rays=n)
lcfg = O.LossConfig()
params = O.init_params(ocfg, seed=0)
rays, gt = H.make_rays(n, seed=1)
eng = Engine(ecfg, H.basis_np())
flat = eng.flatten_params(params); eng.params_changed(flat)
dev = flat.device
rays = {k: v.to(dev) for k, v in rays.items()}; gt = gt.to(dev)
jit = torch.rand(2, n, device=dev)
mu, nu = torch.zeros_like(flat), torch.zeros_like(flat)
lc = _loss_cfg(lcfg)
grad = torch.empty_like(flat); stats = torch.empty(16, device=dev)
a = _lib.AdamCfg(); a.lr, a.beta1, a.beta2, a.eps, a.grad_max_norm, a.grad_max_val, a.step, a.grad_scale = 2e-3, .9, .999, 1e-6, 1e-3, 0., 0, 1.0

def timeit(fn, it=10, warm=3):
  for _ in range(warm): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(it): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / it

t_fwd = timeit(lambda: eng.forward(flat, rays, 0.5, None, compute_extras=True, want_history=False))
t_lg = timeit(lambda: eng.loss_and_grad(flat, rays, gt, 0.5, jit, lc, grad, stats))
t_adam = timeit(lambda: eng.adam_step(flat, grad, mu, nu, a))
fl_fwd = n * 251.4e6
print(f'rays {n}: render fwd {t_fwd:.3f} ms ({n / t_fwd * 1e3:.3e} rays/s, {fl_fwd / t_fwd / 1e9:.1f} TFLOP/s) | '
      f'loss+grad {t_lg:.3f} ms | adam {t_adam:.3f} ms | train step {t_lg + t_adam:.3f} ms '
      f'({n / (t_lg + t_adam) * 1e3:.3e} rays/s, {3 * fl_fwd / (t_lg + t_adam) / 1e9:.1f} TFLOP/s)')

def classes(fn, it=10):
  eng.profile(True)
  for _ in range(it): fn()
  torch.cuda.synchronize()
  prof = eng.profile_read(); eng.profile(False)
  return {k: round(v[0] / it, 3) for k, v in prof.items() if v[1] > 0}
print('render classes ms:', classes(lambda: eng.forward(flat, rays, 0.5, None, compute_extras=True, want_history=False)))
print('train  classes ms:', classes(lambda: eng.loss_and_grad(flat, rays, gt, 0.5, jit, lc, grad, stats)))
