"""CPU oracle: a restatement of the reference's Mip-NeRF 360 per-ray hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`nerf_hugs_b200`) never does: it fails
loudly when its CUDA library is missing.

What it restates (all paths relative to /root/reference/MipNeRF360/internal):
  stepfun.py   :30-128, :131-276, :298-308
  math.py      :26-38, :66-98, :108-127
  coord.py     :21-36, :39-60 (closed-form Jacobian, see `contract_jacobian_apply`),
               :63-99, :102-145
  render.py    :21-41, :44-78, :81-100, :103-151, :185-244
  models.py    :131-330 (Model.__call__), :405-550 (MLP.__call__)
  train_utils.py :72-111, :228-248, :351-369, :487-512

Why a restatement: the reference path is JAX/Flax and jax, jaxlib, flax, optax
and gin are not installable in this environment.  Pinning:
  * the reference's own torch twins of the sampling / loss algebra
    (`nerfacto/utils/ray_utils.py`, `nerfacto/utils/loss_utils.py`) and its pure
    NumPy `geopoly.py` ARE importable; `tests/golden/make_golden.py` ran them
    here and committed their outputs as fixtures, and `tests/test_oracle_golden.py`
    checks this file against them;
  * the RNG-free known-answer tests of `MipNeRF360/tests/*` are re-asserted in
    `tests/test_oracle_reference_properties.py`.
  The MLP / flax.Dense / optax.adam arithmetic has no golden vector anywhere in
  the reference ("parity unpinned" for those pieces); it is
  restated from the published definitions.

Numerics: everything runs in `dtype` (float32 to mirror the reference on CPU,
float64 as the tie-breaking twin).  `quant='bf16'` rounds the *inputs and
kernels* of every Dense layer to bfloat16 (fp32 accumulate, fp32 bias) which is
the arithmetic the tensor-core throughput mode of the CUDA path performs;
`quant='bf16_train'` additionally rounds every Dense layer's incoming gradient
to bfloat16 in the backward pass, which is what that mode's saved dZ tiles do.
"""
from __future__ import annotations

import dataclasses
import math as pymath
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

F32_EPS = float(np.finfo(np.float32).eps)  # jnp.finfo(jnp.float32).eps
PI = pymath.pi


# ----------------------------------------------------------------------------
# math.py
# ----------------------------------------------------------------------------
def safe_sin(x: torch.Tensor) -> torch.Tensor:
  """math.py:26-38: sin(where(|x| < t, x, x % t)), t = 100*pi in x's dtype.

  `%` is Python/NumPy remainder (result has the sign of the divisor) — torch.remainder.
  """
  t = torch.tensor(100.0 * PI, dtype=x.dtype)
  return torch.sin(torch.where(x.abs() < t, x, torch.remainder(x, t)))


def sorted_interp(x, xp, fp, chunk: int = 512):
  """math.py:108-127, brute-force inverse-CDF lookup (chunked over rays only)."""
  outs = []
  lead = x.shape[:-1]
  x2 = x.reshape(-1, x.shape[-1])
  xp2 = xp.expand(lead + xp.shape[-1:]).reshape(-1, xp.shape[-1])
  fp2 = fp.expand(lead + fp.shape[-1:]).reshape(-1, fp.shape[-1])
  for i in range(0, x2.shape[0], chunk):
    xc, xpc, fpc = x2[i:i + chunk], xp2[i:i + chunk], fp2[i:i + chunk]
    mask = xc[:, None, :] >= xpc[:, :, None]

    def find_interval(v):
      v0 = torch.where(mask, v[:, :, None], v[:, :1, None]).amax(-2)
      v1 = torch.where(~mask, v[:, :, None], v[:, -1:, None]).amin(-2)
      return v0, v1

    fp0, fp1 = find_interval(fpc)
    xp0, xp1 = find_interval(xpc)
    offset = torch.clip(torch.nan_to_num((xc - xp0) / (xp1 - xp0), nan=0.0), 0, 1)
    outs.append(fp0 + offset * (fp1 - fp0))
  return torch.cat(outs, 0).reshape(x.shape)


def sorted_interp_index(x, xp):
  """Index i = max{k : x >= xp[k]} that `sorted_interp` implicitly selects (0 if none)."""
  mask = x[..., None, :] >= xp[..., :, None]
  i = torch.arange(xp.shape[-1])
  return torch.where(mask, i[:, None], i[:1, None]).amax(-2)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
  """math.py:66-98 (log_lerp :56-63)."""
  if lr_delay_steps > 0:
    delay_rate = lr_delay_mult + (1 - lr_delay_mult) * pymath.sin(
        0.5 * PI * min(max(step / lr_delay_steps, 0.0), 1.0))
  else:
    delay_rate = 1.0
  t = min(max(step / max_steps, 0.0), 1.0)
  lv0, lv1 = pymath.log(lr_init), pymath.log(lr_final)
  return delay_rate * pymath.exp(t * (lv1 - lv0) + lv0)


# ----------------------------------------------------------------------------
# stepfun.py
# ----------------------------------------------------------------------------
def searchsorted(a, v):
  """stepfun.py:30-54."""
  i = torch.arange(a.shape[-1])
  v_ge_a = v[..., None, :] >= a[..., :, None]
  idx_lo = torch.where(v_ge_a, i[:, None], i[:1, None]).amax(-2)
  idx_hi = torch.where(~v_ge_a, i[:, None], i[-1:, None]).amin(-2)
  return idx_lo, idx_hi


def inner_outer(t0, t1, y1):
  """stepfun.py:64-77."""
  cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, -1)], -1)
  idx_lo, idx_hi = searchsorted(t1, t0)
  cy1_lo = torch.gather(cy1, -1, idx_lo)
  cy1_hi = torch.gather(cy1, -1, idx_hi)
  y0_outer = cy1_hi[..., 1:] - cy1_lo[..., :-1]
  y0_inner = torch.where(idx_hi[..., :-1] <= idx_lo[..., 1:],
                         cy1_lo[..., 1:] - cy1_hi[..., :-1], torch.zeros_like(y0_outer))
  return y0_inner, y0_outer


def lossfun_outer(t, w, t_env, w_env, eps=F32_EPS):
  """stepfun.py:80-86."""
  _, w_outer = inner_outer(t, t_env, w_env)
  return torch.clamp_min(w - w_outer, 0) ** 2 / (w + eps)


def weight_to_pdf(t, w, eps=F32_EPS ** 2):
  """stepfun.py:89-91."""
  return w / torch.clamp_min(t[..., 1:] - t[..., :-1], eps)


def pdf_to_weight(t, p):
  """stepfun.py:94-96."""
  return p * (t[..., 1:] - t[..., :-1])


def max_dilate(t, w, dilation, domain=(-float('inf'), float('inf'))):
  """stepfun.py:99-113."""
  t0 = t[..., :-1] - dilation
  t1 = t[..., 1:] + dilation
  t_dilate = torch.sort(torch.cat([t, t0, t1], -1), -1).values
  t_dilate = torch.clip(t_dilate, domain[0], domain[1])
  inside = (t0[..., None, :] <= t_dilate[..., None]) & (t1[..., None, :] > t_dilate[..., None])
  w_dilate = torch.where(inside, w[..., None, :], torch.zeros((), dtype=w.dtype)).amax(-1)[..., :-1]
  return t_dilate, w_dilate


def max_dilate_weights(t, w, dilation, domain=(-float('inf'), float('inf')),
                       renormalize=False, eps=F32_EPS ** 2):
  """stepfun.py:116-128."""
  p = weight_to_pdf(t, w)
  t_dilate, p_dilate = max_dilate(t, p, dilation, domain=domain)
  w_dilate = pdf_to_weight(t_dilate, p_dilate)
  if renormalize:
    w_dilate = w_dilate / torch.clamp_min(w_dilate.sum(-1, keepdim=True), eps)
  return t_dilate, w_dilate


def integrate_weights(w):
  """stepfun.py:131-150."""
  cw = torch.clamp_max(torch.cumsum(w[..., :-1], -1), 1)
  shape = cw.shape[:-1] + (1,)
  return torch.cat([torch.zeros(shape, dtype=w.dtype), cw, torch.ones(shape, dtype=w.dtype)], -1)


def invert_cdf(u, t, w_logits):
  """stepfun.py:153-161 (use_gpu_resampling=False branch -> sorted_interp)."""
  w = torch.softmax(w_logits, -1)
  cw = integrate_weights(w)
  return sorted_interp(u, cw, t)


def sample_u(num_samples: int, deterministic_center: bool, jitter: Optional[torch.Tensor],
             dtype=torch.float32) -> Tuple[torch.Tensor, float]:
  """The `u` construction of stepfun.py:188-209.

  Returns (u_base[num_samples], max_jitter).  `jitter` is the caller-supplied
  per-ray uniform draw in [0, 1) (the reference draws `jax.random.uniform(...,
  maxval=max_jitter)`; threefry cannot be reproduced without JAX so the draw is
  an input and is scaled by max_jitter here: u = u_base + jitter * max_jitter).
  linspace endpoints are computed in float64 and rounded once to `dtype`
  (canonical choice; jnp.linspace's ulps are not reproducible without JAX).
  """
  eps = F32_EPS
  if jitter is None:
    if deterministic_center:
      pad = 1 / (2 * num_samples)
      u = np.linspace(pad, 1. - pad - eps, num_samples)
    else:
      u = np.linspace(0, 1. - eps, num_samples)
    return torch.tensor(u, dtype=dtype), 0.0
  u_max = eps + (1 - eps) / num_samples
  max_jitter = (1 - u_max) / (num_samples - 1) - eps
  u = np.linspace(0, 1 - u_max, num_samples)
  return torch.tensor(u, dtype=dtype), float(max_jitter)


def sample(jitter, t, w_logits, num_samples, single_jitter=False, deterministic_center=False):
  """stepfun.py:164-211.  `jitter`: None (rng=None) or tensor [..., 1|num_samples] in [0,1)."""
  u_base, max_jitter = sample_u(num_samples, deterministic_center, jitter, t.dtype)
  if jitter is None:
    u = u_base.expand(t.shape[:-1] + (num_samples,))
  else:
    u = u_base + jitter.to(t.dtype) * torch.tensor(max_jitter, dtype=t.dtype)
  return invert_cdf(u, t, w_logits)


def sample_intervals(jitter, t, w_logits, num_samples, single_jitter=False,
                     domain=(-float('inf'), float('inf'))):
  """stepfun.py:214-263."""
  if num_samples <= 1:
    raise ValueError(f'num_samples must be > 1, is {num_samples}.')
  centers = sample(jitter, t, w_logits, num_samples, single_jitter, deterministic_center=True)
  mid = (centers[..., 1:] + centers[..., :-1]) / 2
  minval, maxval = domain
  first = torch.clamp_min(2 * centers[..., :1] - mid[..., :1], minval)
  last = torch.clamp_max(2 * centers[..., -1:] - mid[..., -1:], maxval)
  return torch.cat([first, mid, last], -1)


def lossfun_distortion(t, w):
  """stepfun.py:266-276 (the O(S^2) definition)."""
  ut = (t[..., 1:] + t[..., :-1]) / 2
  dut = (ut[..., :, None] - ut[..., None, :]).abs()
  loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, -1), -1)
  loss_intra = torch.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), -1) / 3
  return loss_inter + loss_intra


def _interp_1d(x, xp, fp):
  """jnp.interp for sorted xp with batch dims: x [...,m], xp/fp [...,n]."""
  n = xp.shape[-1]
  idx = torch.searchsorted(xp.contiguous(), x.contiguous(), right=True)
  i0 = torch.clamp(idx - 1, 0, n - 1)
  i1 = torch.clamp(idx, 0, n - 1)
  x0, x1 = torch.gather(xp, -1, i0), torch.gather(xp, -1, i1)
  f0, f1 = torch.gather(fp, -1, i0), torch.gather(fp, -1, i1)
  dx = x1 - x0
  # jnp.interp: f = fp[i-1] + (x - xp[i-1]) / dx * df, with dx==0 -> fp[i-1]-ish; clamp ends.
  # jax's _interp: dx0 = |dx| <= np.spacing(finfo(dtype).eps)  (1.42e-14 in float32) -> fp[i-1]
  tiny = float(np.spacing(np.finfo(np.float32 if x.dtype == torch.float32 else np.float64).eps))
  dx0 = dx.abs() <= tiny
  w = torch.where(dx0, torch.zeros_like(dx), (x - x0) / torch.where(dx0, torch.ones_like(dx), dx))
  out = f0 + w * (f1 - f0)
  out = torch.where(x < xp[..., :1], fp[..., :1].expand_as(out), out)
  out = torch.where(x > xp[..., -1:], fp[..., -1:].expand_as(out), out)
  return out


def weighted_percentile(t, w, ps):
  """stepfun.py:298-308."""
  cw = integrate_weights(w)
  q = torch.tensor(ps, dtype=t.dtype) / 100
  return _interp_1d(q.expand(t.shape[:-1] + (len(ps),)), cw, t)


# ----------------------------------------------------------------------------
# coord.py
# ----------------------------------------------------------------------------
def contract(x):
  """coord.py:21-27."""
  x_mag_sq = torch.clamp_min(torch.sum(x ** 2, -1, keepdim=True), F32_EPS)
  return torch.where(x_mag_sq <= 1, x, ((2 * torch.sqrt(x_mag_sq) - 1) / x_mag_sq) * x)


def contract_jacobian_apply(x, v):
  """J(x) @ v for the contraction, closed form.

  Replaces jax.linearize in coord.py:58-59.  m = max(eps, |x|^2).  For m <= 1
  J = I.  Else z = s x with s = (2 sqrt(m) - 1)/m and
  J = s I + x (ds/dx)^T,  ds/dx = 2 (m^-2 - m^-3/2) x.   x: [...,3], v: [...,3].
  """
  m = torch.clamp_min(torch.sum(x ** 2, -1, keepdim=True), F32_EPS)
  sq = torch.sqrt(m)
  s = (2 * sq - 1) / m
  c = 2 * (1 / (m * m) - 1 / (m * sq))
  xv = torch.sum(x * v, -1, keepdim=True)
  return torch.where(m <= 1, v, s * v + c * xv * x)


def track_linearize_contract(mean, cov):
  """coord.py:39-60 specialised to fn=contract: (contract(mean), J cov J^T)."""
  fn_mean = contract(mean)
  # cov [...,3,3]; apply J to columns then rows (J symmetric).
  x = mean[..., None, :]
  jc = contract_jacobian_apply(x, cov.transpose(-1, -2)).transpose(-1, -2)   # J @ cov
  fn_cov = contract_jacobian_apply(x, jc)                                     # (J cov) J^T, row-wise
  return fn_mean, fn_cov


def construct_ray_warps(fn: Optional[str], t_near, t_far):
  """coord.py:63-99.  fn in {None, 'reciprocal', 'log', 'piecewise'}."""
  if fn is None:
    fwd = inv = lambda x: x
  elif fn == 'piecewise':
    fwd = lambda x: torch.where(x < 1, .5 * x, 1 - .5 / x)
    inv = lambda x: torch.where(x < .5, 2 * x, .5 / (1 - x))
  elif fn == 'reciprocal':
    fwd = inv = torch.reciprocal
  elif fn == 'log':
    fwd, inv = torch.log, torch.exp
  else:
    raise ValueError(fn)
  s_near, s_far = fwd(t_near), fwd(t_far)
  t_to_s = lambda t: (fwd(t) - s_near) / (s_far - s_near)
  s_to_t = lambda s: inv(s * s_far + (1 - s) * s_near)
  return t_to_s, s_to_t


def integrated_pos_enc(mean, var, min_deg, max_deg):
  """coord.py:102-126.  Output column order: [deg-major sin | deg-major sin(.+pi/2)]."""
  scales = torch.tensor([2.0 ** k for k in range(min_deg, max_deg)], dtype=mean.dtype)
  shape = mean.shape[:-1] + (-1,)
  scaled_mean = (mean[..., None, :] * scales[:, None]).reshape(shape)
  scaled_var = (var[..., None, :] * scales[:, None] ** 2).reshape(shape)
  half_pi = torch.tensor(0.5 * PI, dtype=mean.dtype)
  x = torch.cat([scaled_mean, scaled_mean + half_pi], -1)
  v = torch.cat([scaled_var] * 2, -1)
  return torch.exp(-0.5 * v) * safe_sin(x)


def lift_and_diagonalize(mean, cov, basis):
  """coord.py:129-133.  basis [3, nb]."""
  fn_mean = mean @ basis
  fn_cov_diag = torch.sum(basis * (cov @ basis), -2)
  return fn_mean, fn_cov_diag


def pos_enc(x, min_deg, max_deg, append_identity=True):
  """coord.py:136-147 (plain sin, no safe_sin)."""
  scales = torch.tensor([2.0 ** k for k in range(min_deg, max_deg)], dtype=x.dtype)
  shape = x.shape[:-1] + (-1,)
  scaled_x = (x[..., None, :] * scales[:, None]).reshape(shape)
  half_pi = torch.tensor(0.5 * PI, dtype=x.dtype)
  four_feat = torch.sin(torch.cat([scaled_x, scaled_x + half_pi], -1))
  return torch.cat([x, four_feat], -1) if append_identity else four_feat


# ----------------------------------------------------------------------------
# render.py
# ----------------------------------------------------------------------------
def lift_gaussian(d, t_mean, t_var, r_var, diag):
  """render.py:21-41."""
  mean = d[..., None, :] * t_mean[..., None]
  d_mag_sq = torch.clamp_min(torch.sum(d ** 2, -1, keepdim=True), 1e-10)
  if diag:
    d_outer_diag = d ** 2
    null_outer_diag = 1 - d_outer_diag / d_mag_sq
    cov_diag = (t_var[..., None] * d_outer_diag[..., None, :] +
                r_var[..., None] * null_outer_diag[..., None, :])
    return mean, cov_diag
  d_outer = d[..., :, None] * d[..., None, :]
  eye = torch.eye(d.shape[-1], dtype=d.dtype)
  null_outer = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
  t_cov = t_var[..., None, None] * d_outer[..., None, :, :]
  xy_cov = r_var[..., None, None] * null_outer[..., None, :, :]
  return mean, t_cov + xy_cov


def conical_frustum_to_gaussian(d, t0, t1, base_radius, diag):
  """render.py:44-78 (stable=True branch)."""
  mu = (t0 + t1) / 2
  hw = (t1 - t0) / 2
  eps = F32_EPS
  t_mean = mu + (2 * mu * hw ** 2) / torch.clamp_min(3 * mu ** 2 + hw ** 2, eps)
  denom = torch.clamp_min(3 * mu ** 2 + hw ** 2, eps)
  t_var = (hw ** 2) / 3 - (4 / 15) * hw ** 4 * (12 * mu ** 2 - hw ** 2) / denom ** 2
  r_var = (mu ** 2) / 4 + (5 / 12) * hw ** 2 - (4 / 15) * (hw ** 4) / denom
  r_var = r_var * base_radius ** 2
  return lift_gaussian(d, t_mean, t_var, r_var, diag)


def cylinder_to_gaussian(d, t0, t1, radius, diag):
  """render.py:81-100."""
  t_mean = (t0 + t1) / 2
  r_var = radius ** 2 / 4
  t_var = (t1 - t0) ** 2 / 12
  return lift_gaussian(d, t_mean, t_var, r_var.expand_as(t_mean), diag)


def cast_rays(tdist, origins, directions, radii, ray_shape, diag=True):
  """render.py:103-127."""
  t0, t1 = tdist[..., :-1], tdist[..., 1:]
  fn = {'cone': conical_frustum_to_gaussian, 'cylinder': cylinder_to_gaussian}[ray_shape]
  means, covs = fn(directions, t0, t1, radii, diag)
  return means + origins[..., None, :], covs


def compute_alpha_weights(density, tdist, dirs, opaque_background=False):
  """render.py:130-151."""
  t_delta = tdist[..., 1:] - tdist[..., :-1]
  delta = t_delta * torch.linalg.norm(dirs[..., None, :], dim=-1)
  density_delta = density * delta
  if opaque_background:
    density_delta = torch.cat([density_delta[..., :-1],
                               torch.full_like(density_delta[..., -1:], float('inf'))], -1)
  alpha = 1 - torch.exp(-density_delta)
  trans = torch.exp(-torch.cat([torch.zeros_like(density_delta[..., :1]),
                                torch.cumsum(density_delta[..., :-1], -1)], -1))
  return alpha * trans, alpha, trans


def volumetric_rendering(rgbs, weights, tdist, bg_rgbs, t_far, compute_extras):
  """render.py:185-244 (extras=None)."""
  eps = F32_EPS
  rendering = {}
  acc = weights.sum(-1)
  bg_w = torch.clamp_min(1 - acc[..., None], 0)
  rendering['rgb'] = (weights[..., None] * rgbs).sum(-2) + bg_w * bg_rgbs
  if compute_extras:
    rendering['acc'] = acc
    expectation = lambda x: (weights * x).sum(-1) / torch.clamp_min(acc, eps)
    t_mids = 0.5 * (tdist[..., :-1] + tdist[..., 1:])
    # `jnp.nan_to_num(x, jnp.inf)` (render.py:222): the second positional parameter of nan_to_num is `copy`, not `nan`,
    # so NaN -> 0.0 and +inf -> float32 max (pinned by tests/golden/mip360_ops.npz::vr_nan_distance_mean)
    dm = torch.nan_to_num(torch.exp(expectation(torch.log(t_mids))), nan=0.0)
    rendering['distance_mean'] = torch.minimum(torch.maximum(dm, tdist[..., 0]), tdist[..., -1])
    t_aug = torch.cat([tdist, t_far], -1)
    weights_aug = torch.cat([weights, bg_w], -1)
    ps = [5, 50, 95]
    pct = weighted_percentile(t_aug, weights_aug, ps)
    for i, p in enumerate(ps):
      s = 'median' if p == 50 else 'percentile_' + str(p)
      rendering['distance_' + s] = pct[..., i]
  return rendering


# ----------------------------------------------------------------------------
# models.py
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class MLPConfig:
  """models.py:360-391 (the fields the shipped gins bind)."""
  net_depth: int = 8
  net_width: int = 256
  bottleneck_width: int = 256
  net_depth_viewdirs: int = 1
  net_width_viewdirs: int = 128
  min_deg_point: int = 0
  max_deg_point: int = 12
  skip_layer: int = 4
  num_rgb_channels: int = 3
  deg_view: int = 4
  density_bias: float = -1.
  rgb_premultiplier: float = 1.
  rgb_bias: float = 0.
  rgb_padding: float = 0.001
  disable_rgb: bool = False
  warp_fn: Optional[str] = None      # None | 'contract'


@dataclasses.dataclass
class ModelConfig:
  """models.py:47-71 + the Config fields the hot path reads (configs.py:45-184)."""
  num_prop_samples: int = 64
  num_nerf_samples: int = 32
  num_levels: int = 3
  bg_intensity: float = 1.0
  anneal_slope: float = 10
  use_viewdirs: bool = True
  raydist_fn: Optional[str] = None   # None | 'reciprocal' | ...
  ray_shape: str = 'cone'
  single_jitter: bool = True
  dilation_multiplier: float = 0.5
  dilation_bias: float = 0.0025
  num_glo_features: int = 0
  num_embeddings: int = 3500
  near_anneal_rate: Optional[float] = None
  near_anneal_init: float = 0.95
  resample_padding: float = 0.0
  opaque_background: bool = False
  nerf_mlp: MLPConfig = dataclasses.field(default_factory=MLPConfig)
  prop_mlp: MLPConfig = dataclasses.field(
      default_factory=lambda: MLPConfig(net_depth=4, net_width=256, disable_rgb=True))


def mlp_param_shapes(cfg: MLPConfig, num_glo_features: int, basis_n: int = 21) -> List[Tuple[str, Tuple[int, int]]]:
  """Dense layer (in, out) shapes in flax creation order (Dense_0, Dense_1, ...), models.py:449-515."""
  in_dim = 2 * basis_n * (cfg.max_deg_point - cfg.min_deg_point)
  shapes = []
  d = in_dim
  for i in range(cfg.net_depth):
    shapes.append((d, cfg.net_width))
    d = cfg.net_width
    if i % cfg.skip_layer == 0 and i > 0:
      d = cfg.net_width + in_dim
  shapes.append((d, 1))
  if not cfg.disable_rgb:
    shapes.append((d, cfg.bottleneck_width))
    dv = cfg.bottleneck_width + (3 + 3 * 2 * cfg.deg_view) + num_glo_features
    for i in range(cfg.net_depth_viewdirs):
      shapes.append((dv, cfg.net_width_viewdirs))
      dv = cfg.net_width_viewdirs
    shapes.append((dv, cfg.num_rgb_channels))
  return [(f'Dense_{i}', s) for i, s in enumerate(shapes)]


def init_params(cfg: ModelConfig, seed: int = 0, dtype=torch.float32, bias_scale: float = 0.0):
  """Random-init parameter tree with flax names.

  flax.linen.Dense(kernel_init=he_uniform) -> kernel ~ U(-sqrt(6/fan_in), +), bias = 0
  (models.py:432-433).  `bias_scale` > 0 draws non-zero biases so parity tests exercise
  the bias path too.
  """
  g = torch.Generator().manual_seed(seed)
  params: Dict[str, Dict[str, Dict[str, torch.Tensor]]] = {}
  for name, mcfg in (('NerfMLP_0', cfg.nerf_mlp), ('PropMLP_0', cfg.prop_mlp)):
    glo = cfg.num_glo_features if name == 'NerfMLP_0' else 0
    layers = {}
    for lname, (fi, fo) in mlp_param_shapes(mcfg, glo):
      bound = pymath.sqrt(6.0 / fi)
      k = (torch.rand(fi, fo, generator=g, dtype=torch.float64) * 2 - 1) * bound
      b = (torch.rand(fo, generator=g, dtype=torch.float64) * 2 - 1) * bias_scale
      layers[lname] = {'kernel': k.to(dtype), 'bias': b.to(dtype)}
    params[name] = layers
  if cfg.num_glo_features > 0:
    # flax nn.Embed default init: normal(stddev=1/sqrt(features))... (variance_scaling fan_in, out axis)
    e = torch.randn(cfg.num_embeddings, cfg.num_glo_features, generator=g, dtype=torch.float64)
    params['GloEmbed_0'] = {'embedding': (e / pymath.sqrt(cfg.num_glo_features)).to(dtype)}
  return params


def _q(x, quant):
  return x.to(torch.bfloat16).to(x.dtype) if quant in ('bf16', 'bf16_train') else x


class _DenseBf16Train(torch.autograd.Function):
  """Dense layer as the tensor-core throughput mode trains it: bf16-rounded inputs and kernel in the forward pass
  (fp32 accumulate, fp32 bias) and, in the backward pass, the incoming gradient dZ rounded to bf16 *before* it is
  used - the chain kernel keeps dZ as a bf16 tile that feeds the dgrad MMA, the weight-gradient GEMM and the bias
  column sums alike (csrc/mlp_pp.cu backward epilogues, csrc/wgrad_tc.cu).  The saved activation is the bf16 one."""

  @staticmethod
  def forward(ctx, x, kernel, bias):
    xq, kq = _q(x, 'bf16'), _q(kernel, 'bf16')
    ctx.save_for_backward(xq, kq)
    return xq @ kq + bias

  @staticmethod
  def backward(ctx, g):
    xq, kq = ctx.saved_tensors
    gq = _q(g, 'bf16')
    g2 = gq.reshape(-1, gq.shape[-1])
    return gq @ kq.T, xq.reshape(-1, xq.shape[-1]).T @ g2, g2.sum(0)


def _dense(x, layer, quant):
  if quant == 'bf16_train':
    return _DenseBf16Train.apply(x, layer['kernel'], layer['bias'])
  return _q(x, quant) @ _q(layer['kernel'], quant) + layer['bias']


def mlp_apply(mcfg: MLPConfig, p, means, covs, viewdirs, glo_vec, basis, quant=None,
              return_features=False, features=None):
  """models.py:405-550 (transient head omitted: out of scope).  `features`: test hook, IPE features computed elsewhere
  (e.g. the CUDA encoder's bf16 features) replace the encoding of (means, covs)."""
  if features is not None:
    x = features
  else:
    if mcfg.warp_fn == 'contract':
      means, covs = track_linearize_contract(means, covs)
    elif mcfg.warp_fn is not None:
      raise ValueError(mcfg.warp_fn)
    lifted_means, lifted_vars = lift_and_diagonalize(means, covs, basis)
    x = integrated_pos_enc(lifted_means, lifted_vars, mcfg.min_deg_point, mcfg.max_deg_point)
  inputs = x
  li = 0
  for i in range(mcfg.net_depth):
    x = torch.relu(_dense(x, p[f'Dense_{li}'], quant)); li += 1
    if i % mcfg.skip_layer == 0 and i > 0:
      x = torch.cat([x, inputs], -1)
  raw_density = _dense(x, p[f'Dense_{li}'], quant)[..., 0]; li += 1
  density = torch.nn.functional.softplus(raw_density + mcfg.density_bias)
  out = {'density': density, 'raw_density': raw_density}
  if return_features:
    out['features'] = inputs
  if mcfg.disable_rgb:
    out['rgb'] = torch.zeros_like(means)
    return out
  bottleneck = _dense(x, p[f'Dense_{li}'], quant); li += 1
  xs = [bottleneck]
  dir_enc = pos_enc(viewdirs, 0, mcfg.deg_view, True)
  xs.append(dir_enc[..., None, :].expand(bottleneck.shape[:-1] + (dir_enc.shape[-1],)))
  if glo_vec is not None:
    xs.append(glo_vec[..., None, :].expand(bottleneck.shape[:-1] + glo_vec.shape[-1:]))
  x = torch.cat(xs, -1)
  for i in range(mcfg.net_depth_viewdirs):
    x = torch.relu(_dense(x, p[f'Dense_{li}'], quant)); li += 1
  raw_rgb = _dense(x, p[f'Dense_{li}'], quant)
  rgb = torch.sigmoid(mcfg.rgb_premultiplier * raw_rgb + mcfg.rgb_bias)
  out['rgb'] = rgb * (1 + 2 * mcfg.rgb_padding) - mcfg.rgb_padding
  return out


def model_apply(cfg: ModelConfig, params, rays: Dict[str, torch.Tensor], train_frac: float,
                compute_extras: bool, basis: torch.Tensor, jitter: Optional[Sequence[torch.Tensor]] = None,
                zero_glo: bool = False, quant=None, vis_num_rays: int = 16, features=None):
  """models.py:74-330 Model.__call__.

  rays: dict with origins, directions, viewdirs [...,3]; radii, near, far [...,1];
  embed_idx [...,1] int.  jitter: None (rng=None, deterministic) or a list of
  num_levels tensors [..., 1] of uniform draws in [0,1).
  """
  dtype = rays['origins'].dtype
  if cfg.num_glo_features > 0:
    if not zero_glo:
      glo_vec = params['GloEmbed_0']['embedding'][rays['embed_idx'][..., 0].long()]
    else:
      glo_vec = torch.zeros(rays['origins'].shape[:-1] + (cfg.num_glo_features,), dtype=dtype)
  else:
    glo_vec = None
  _, s_to_t = construct_ray_warps(cfg.raydist_fn, rays['near'], rays['far'])
  if cfg.near_anneal_rate is None:
    init_s_near = 0.
  else:
    init_s_near = float(np.clip(1 - train_frac / cfg.near_anneal_rate, 0, cfg.near_anneal_init))
  init_s_far = 1.
  sdist = torch.cat([torch.full_like(rays['near'], init_s_near),
                     torch.full_like(rays['far'], init_s_far)], -1)
  weights = torch.ones_like(rays['near'])
  prod_num_samples = 1
  ray_history, renderings = [], []
  for i_level in range(cfg.num_levels):
    is_prop = i_level < cfg.num_levels - 1
    num_samples = cfg.num_prop_samples if is_prop else cfg.num_nerf_samples
    dilation = cfg.dilation_bias + cfg.dilation_multiplier * (init_s_far - init_s_near) / prod_num_samples
    prod_num_samples *= num_samples
    use_dilation = cfg.dilation_bias > 0 or cfg.dilation_multiplier > 0
    if i_level > 0 and use_dilation:
      sdist, weights = max_dilate_weights(sdist, weights, dilation,
                                          domain=(init_s_near, init_s_far), renormalize=True)
      sdist = sdist[..., 1:-1]
      weights = weights[..., 1:-1]
    if cfg.anneal_slope > 0:
      bias = lambda x, s: (s * x) / ((s - 1) * x + 1)
      anneal = bias(train_frac, cfg.anneal_slope)
    else:
      anneal = 1.
    logits_resample = torch.where(sdist[..., 1:] > sdist[..., :-1],
                                  anneal * torch.log(weights + cfg.resample_padding),
                                  torch.tensor(-float('inf'), dtype=dtype))
    with torch.no_grad():
      sdist = sample_intervals(None if jitter is None else jitter[i_level], sdist.detach(),
                               logits_resample.detach(), num_samples,
                               single_jitter=cfg.single_jitter, domain=(init_s_near, init_s_far))
    tdist = s_to_t(sdist)
    means, covs = cast_rays(tdist, rays['origins'], rays['directions'], rays['radii'],
                            cfg.ray_shape, diag=False)
    mcfg = cfg.prop_mlp if is_prop else cfg.nerf_mlp
    p = params['PropMLP_0' if is_prop else 'NerfMLP_0']
    ray_results = mlp_apply(mcfg, p, means, covs,
                            rays['viewdirs'] if cfg.use_viewdirs else None,
                            None if is_prop else glo_vec, basis, quant=quant,
                            features=None if features is None else features[i_level])
    weights = compute_alpha_weights(ray_results['density'], tdist, rays['directions'],
                                    opaque_background=cfg.opaque_background)[0]
    rendering = volumetric_rendering(ray_results['rgb'], weights, tdist, cfg.bg_intensity,
                                     rays['far'], compute_extras)
    if compute_extras:
      n = vis_num_rays
      rendering['ray_sdist'] = sdist.reshape(-1, sdist.shape[-1])[:n]
      rendering['ray_weights'] = weights.reshape(-1, weights.shape[-1])[:n]
      rgb = ray_results['rgb']
      rendering['ray_rgbs'] = rgb.reshape((-1,) + rgb.shape[-2:])[:n]
    renderings.append(rendering)
    ray_results['sdist'] = sdist.clone()
    ray_results['tdist'] = tdist
    ray_results['weights'] = weights.clone()
    ray_history.append(ray_results)
  if compute_extras:
    ws = [r['ray_weights'] for r in renderings]
    rgbs = [r['ray_rgbs'] for r in renderings]
    final_rgb = torch.sum(rgbs[-1] * ws[-1][..., None], -2)
    for i in range(len(renderings) - 1):
      renderings[i]['ray_rgbs'] = final_rgb[:, None, :].expand(rgbs[i].shape)
  return renderings, ray_history


# ----------------------------------------------------------------------------
# train_utils.py
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class LossConfig:
  """configs.py:84-107,131-132."""
  data_loss_type: str = 'charb'
  charb_padding: float = 0.001
  data_loss_mult: float = 1.0
  data_coarse_loss_mult: float = 0.
  interlevel_loss_mult: float = 1.0
  distortion_loss_mult: float = 0.01
  transient_type: Optional[str] = None   # None | 'withmask'
  withmask_transient_weight: float = 0.
  disable_multiscale_loss: bool = False
  lr_init: float = 0.002
  lr_final: float = 0.00002
  lr_delay_steps: int = 512
  lr_delay_mult: float = 0.01
  max_steps: int = 250000
  adam_beta1: float = 0.9
  adam_beta2: float = 0.999
  adam_eps: float = 1e-6
  grad_max_norm: float = 0.001
  grad_max_val: float = 0.


def compute_data_loss(rgb_gt, rays, renderings, lcfg: LossConfig, use_static_mask: bool):
  """train_utils.py:72-111, including quirk B1 (withmask denom counts rays, not channels)."""
  data_losses, mses = [], []
  static_mask = (rays['static_mask'] >= 0.5).to(rgb_gt.dtype)
  for rendering in renderings:
    if use_static_mask:
      lossmult = static_mask + (1 - static_mask) * lcfg.withmask_transient_weight   # [...,1]
    else:
      lossmult = rays['lossmult'].expand(rgb_gt[..., :3].shape)
      if lcfg.disable_multiscale_loss:
        lossmult = torch.ones_like(lossmult)
    resid_sq = (rendering['rgb'] - rgb_gt[..., :3]) ** 2
    denom = torch.clamp_min(lossmult.sum(), F32_EPS)
    mses.append((lossmult * resid_sq).sum() / denom)
    if lcfg.data_loss_type == 'mse':
      data_loss = resid_sq
    elif lcfg.data_loss_type == 'charb':
      data_loss = torch.sqrt(resid_sq + lcfg.charb_padding ** 2)
    else:
      raise ValueError(lcfg.data_loss_type)
    data_losses.append((lossmult * data_loss).sum() / denom)
  data_losses = torch.stack(data_losses)
  loss = lcfg.data_coarse_loss_mult * data_losses[:-1].sum() + lcfg.data_loss_mult * data_losses[-1]
  return loss, {'mses': torch.stack(mses)}


def interlevel_loss(ray_history, lcfg: LossConfig):
  """train_utils.py:228-239."""
  c = ray_history[-1]['sdist'].detach()
  w = ray_history[-1]['weights'].detach()
  loss = 0.
  for rr in ray_history[:-1]:
    loss = loss + torch.mean(lossfun_outer(c, w, rr['sdist'], rr['weights']))
  return lcfg.interlevel_loss_mult * loss


def distortion_loss(ray_history, lcfg: LossConfig):
  """train_utils.py:242-248."""
  c = ray_history[-1]['sdist']
  w = ray_history[-1]['weights']
  return lcfg.distortion_loss_mult * torch.mean(lossfun_distortion(c, w))


def loss_fn(cfg: ModelConfig, lcfg: LossConfig, params, rays, rgb_gt, train_frac, basis,
            jitter=None, quant=None, features=None):
  """train_utils.py:413-448 (transient_type in {None,'withmask'}, no weight decay)."""
  renderings, ray_history = model_apply(cfg, params, rays, train_frac, False, basis,
                                        jitter=jitter, quant=quant, features=features)
  losses = {}
  losses['data'], stats = compute_data_loss(rgb_gt, rays, renderings, lcfg,
                                            lcfg.transient_type == 'withmask')
  if lcfg.interlevel_loss_mult > 0:
    losses['interlevel'] = interlevel_loss(ray_history, lcfg)
  if lcfg.distortion_loss_mult > 0:
    losses['distortion'] = distortion_loss(ray_history, lcfg)
  stats['losses'] = losses
  stats['loss'] = sum(losses.values())
  return stats['loss'], stats, renderings, ray_history


def tree_leaves(tree, prefix=''):
  out = []
  for k in tree:
    v = tree[k]
    if isinstance(v, dict):
      out += tree_leaves(v, prefix + k + '/')
    else:
      out.append((prefix + k, v))
  return out


def clip_gradients(grads, lcfg: LossConfig):
  """train_utils.py:351-369: per top-level module value clip then norm clip."""
  out = {}
  for k, g in grads.items():
    leaves = dict(tree_leaves(g))
    if lcfg.grad_max_val > 0:
      leaves = {n: torch.clip(z, -lcfg.grad_max_val, lcfg.grad_max_val) for n, z in leaves.items()}
    if lcfg.grad_max_norm > 0:
      norm = torch.sqrt(sum((z ** 2).sum() for z in leaves.values()))
      mult = torch.clamp_max(lcfg.grad_max_norm / (F32_EPS + norm), 1.0)
      leaves = {n: mult * z for n, z in leaves.items()}
    out[k] = leaves
  return out


def train_step(cfg: ModelConfig, lcfg: LossConfig, params, opt_state, step: int, rays, rgb_gt,
               train_frac, basis, jitter=None, quant=None, features=None):
  """train_utils.py:386-477 on one device: value_and_grad -> clip -> nan_to_num -> optax.adam.

  params: nested dict of leaf tensors (updated functionally; new dict returned).
  opt_state: {'mu': tree, 'nu': tree} keyed 'Module/Dense_k/kernel'.
  step: optax count before this update (0-based).
  """
  flat = tree_leaves(params)
  leaves = [v.detach().clone().requires_grad_(True) for _, v in flat]

  def rebuild(vals):
    tree = {}
    for (name, _), v in zip(flat, vals):
      parts = name.split('/')
      d = tree
      for q in parts[:-1]:
        d = d.setdefault(q, {})
      d[parts[-1]] = v
    return tree

  p = rebuild(leaves)
  loss, stats, _, _ = loss_fn(cfg, lcfg, p, rays, rgb_gt, train_frac, basis, jitter, quant, features)
  gl = torch.autograd.grad(loss, leaves, allow_unused=True)
  gl = [torch.zeros_like(l) if g is None else g for g, l in zip(gl, leaves)]
  gtree = rebuild(gl)
  raw_grads = {n: g for (n, _), g in zip(flat, gl)}
  clipped = clip_gradients(gtree, lcfg)
  lr = learning_rate_decay(step, lcfg.lr_init, lcfg.lr_final, lcfg.max_steps,
                           lcfg.lr_delay_steps, lcfg.lr_delay_mult)
  b1, b2, eps = lcfg.adam_beta1, lcfg.adam_beta2, lcfg.adam_eps
  t = step + 1
  new_vals, new_mu, new_nu = [], {}, {}
  for (name, v) in flat:
    mod, rest = name.split('/', 1)
    g = torch.nan_to_num(clipped[mod][rest])
    mu = b1 * opt_state['mu'][name] + (1 - b1) * g
    nu = b2 * opt_state['nu'][name] + (1 - b2) * g * g
    mhat = mu / (1 - b1 ** t)
    vhat = nu / (1 - b2 ** t)
    new_vals.append(v.detach() - lr * mhat / (torch.sqrt(vhat) + eps))
    new_mu[name], new_nu[name] = mu, nu
  stats = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in stats.items()}
  stats['lr'] = lr
  return rebuild(new_vals), {'mu': new_mu, 'nu': new_nu}, stats, raw_grads


def init_opt_state(params):
  flat = tree_leaves(params)
  return {'mu': {n: torch.zeros_like(v) for n, v in flat},
          'nu': {n: torch.zeros_like(v) for n, v in flat}}
