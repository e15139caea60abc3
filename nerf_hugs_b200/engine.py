"""Thin Python host over the C ABI (include/hugs_b200.h): owns the handle, marshals torch
device tensors to raw pointers, and exposes the flat fp32 parameter buffer with flax names.

PyTorch is used for device memory, streams and torch.distributed only; all numerics run in
libhugs_b200.so.  There is no CPU path: constructing an Engine without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

RAYDIST = {None: 0, 'reciprocal': 1, 'log': 2, 'piecewise': 3}
RAY_SHAPE = {'cone': 0, 'cylinder': 1}
PRECISION = {'fp32': 0, 'bf16_tc': 1, 'tc_split': 2}
ENCODING = {'ipe': 0, 'point_pe': 1}


@dataclasses.dataclass
class EngineConfig:
  """Mirror of hugs_model_desc (the gin-bound fields of models.Model / NerfMLP / PropMLP)."""
  num_levels: int = 3
  num_prop_samples: int = 64
  num_nerf_samples: int = 32
  nerf_depth: int = 8
  nerf_width: int = 256
  prop_depth: int = 4
  prop_width: int = 256
  bottleneck_width: int = 256
  view_width: int = 128
  skip_layer: int = 4
  min_deg_point: int = 0
  max_deg_point: int = 12
  deg_view: int = 4
  raydist_fn: Optional[str] = None
  ray_shape: str = 'cone'
  nerf_contract: bool = False
  prop_contract: bool = False
  opaque_background: bool = False
  bg_intensity: float = 1.0
  anneal_slope: float = 10.0
  dilation_multiplier: float = 0.5
  dilation_bias: float = 0.0025
  resample_padding: float = 0.0
  near_anneal_rate: Optional[float] = None
  near_anneal_init: float = 0.95
  num_glo_features: int = 0
  num_embeddings: int = 3500
  density_bias: float = -1.0
  rgb_premultiplier: float = 1.0
  rgb_bias: float = 0.0
  rgb_padding: float = 0.001
  precision: str = 'bf16_tc'
  max_rays: int = 4096
  encoding: str = 'ipe'           # 'point_pe': the torch twin's pos_enc of interval midpoints (nerfacto/models/nerf.py)


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, device):
  return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
  """One model instance bound to one CUDA device (one process per GPU)."""

  def __init__(self, cfg: EngineConfig, basis: np.ndarray, device: Optional[torch.device] = None):
    if not torch.cuda.is_available():
      raise RuntimeError('nerf_hugs_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    self.cfg = cfg
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    basis = np.asarray(basis, dtype=np.float32)            # [3, nb] == MLP.pos_basis_t
    assert basis.ndim == 2 and basis.shape[0] == 3 and basis.shape[1] <= 32
    d = _lib.ModelDesc()
    for f in ('num_levels', 'num_prop_samples', 'num_nerf_samples', 'nerf_depth', 'nerf_width', 'prop_depth',
              'prop_width', 'bottleneck_width', 'view_width', 'skip_layer', 'min_deg_point', 'max_deg_point',
              'deg_view', 'bg_intensity', 'anneal_slope', 'dilation_multiplier', 'dilation_bias',
              'resample_padding', 'near_anneal_init', 'num_glo_features', 'num_embeddings', 'density_bias',
              'rgb_premultiplier', 'rgb_bias', 'rgb_padding', 'max_rays'):
      setattr(d, f, getattr(cfg, f))
    d.num_basis = basis.shape[1]
    self.num_basis = int(basis.shape[1])
    flat = np.zeros(96, np.float32)
    flat[:basis.size] = basis.reshape(-1)
    d.basis = (C.c_float * 96)(*flat.tolist())
    d.raydist_fn = RAYDIST[cfg.raydist_fn]
    d.ray_shape = RAY_SHAPE[cfg.ray_shape]
    d.nerf_contract, d.prop_contract = int(cfg.nerf_contract), int(cfg.prop_contract)
    d.opaque_background = int(cfg.opaque_background)
    d.near_anneal_rate = -1.0 if cfg.near_anneal_rate is None else cfg.near_anneal_rate
    d.precision = PRECISION[cfg.precision]
    d.encoding = ENCODING[cfg.encoding]
    self._h = C.c_void_p()
    with torch.cuda.device(self.device):
      check(lib.hugs_create(C.byref(d), C.byref(self._h)))
    self.n_params = int(lib.hugs_param_count(self._h))
    cnt = C.c_int32()
    check(lib.hugs_param_layout(self._h, None, 0, C.byref(cnt)))
    arr = (_lib.TensorDesc * cnt.value)()
    check(lib.hugs_param_layout(self._h, arr, cnt.value, C.byref(cnt)))
    self.layout = [(t.name.decode(), int(t.offset), int(t.rows), int(t.cols), int(t.module)) for t in arr]
    self._keep: List[torch.Tensor] = []

  def close(self):
    if self._h:
      lib.hugs_destroy(self._h)
      self._h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  # ---- parameters ---------------------------------------------------------------------------
  def level_samples(self, level):
    return self.cfg.num_prop_samples if level < self.cfg.num_levels - 1 else self.cfg.num_nerf_samples

  def flatten_params(self, tree) -> torch.Tensor:
    """Nested flax-style dict {Module: {Dense_k: {kernel,bias}}} -> flat fp32 device buffer."""
    flat = torch.zeros(self.n_params, dtype=torch.float32)
    for name, off, rows, cols, _ in self.layout:
      node = tree
      for part in name.split('/'):
        node = node[part]
      flat[off:off + rows * cols] = torch.as_tensor(np.asarray(node), dtype=torch.float32).reshape(-1)
    return flat.to(self.device)

  def unflatten_params(self, flat: torch.Tensor):
    tree: Dict = {}
    host = flat.detach().cpu()
    for name, off, rows, cols, _ in self.layout:
      parts = name.split('/')
      node = tree
      for q in parts[:-1]:
        node = node.setdefault(q, {})
      v = host[off:off + rows * cols]
      node[parts[-1]] = v.reshape(cols) if parts[-1] == 'bias' else v.reshape(rows, cols)
    return tree

  def params_changed(self, params: torch.Tensor):
    check(lib.hugs_params_changed(self._h, _ptr(params), self._stream()))

  # ---- helpers ------------------------------------------------------------------------------
  def _stream(self):
    return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  def _rays(self, rays: Dict[str, torch.Tensor]):
    r = _lib.Rays()
    keep = []
    n = rays['origins'].reshape(-1, 3).shape[0]
    for k in ('origins', 'directions', 'viewdirs', 'radii', 'near', 'far', 'lossmult', 'static_mask'):
      if rays.get(k) is not None:
        t = _f32(rays[k], self.device).reshape(n, -1)
        keep.append(t)
        setattr(r, k, t.data_ptr())
    if rays.get('embed_idx') is not None:
      t = rays['embed_idx'].to(device=self.device, dtype=torch.int32).contiguous().reshape(n, -1)
      keep.append(t)
      r.embed_idx = t.data_ptr()
    return r, keep, n

  def check_embed_idx(self, embed_idx):
    """Raises if a GLO row index is outside [0, num_embeddings) (nn.Embed / jnp.take never touch memory out of range; the
    kernels clamp, this makes the mistake loud).  Synchronises: call it when a dataset is set up, not per step."""
    if self.cfg.num_glo_features > 0 and embed_idx is not None:
      t = torch.as_tensor(embed_idx)
      lo, hi = int(t.min()), int(t.max())
      if lo < 0 or hi >= self.cfg.num_embeddings:
        raise IndexError(f'embed_idx range [{lo}, {hi}] is outside the GLO table of {self.cfg.num_embeddings} rows '
                         '(Model.num_embeddings)')

  # ---- model-level calls --------------------------------------------------------------------
  def forward(self, params: torch.Tensor, rays: Dict[str, torch.Tensor], train_frac: float,
              jitter: Optional[torch.Tensor] = None, compute_extras: bool = True, zero_glo: bool = False,
              want_history: bool = True):
    """Model.__call__ (models.py:74-330): returns (renderings, ray_history) lists of dicts of device tensors."""
    r, keep, n = self._rays(rays)
    L = self.cfg.num_levels
    outs = (_lib.LevelOut * L)()
    res, hist = [], []
    dev = self.device
    for l in range(L):
      S = self.level_samples(l)
      rd = {'rgb': torch.empty(n, 3, device=dev)}
      if compute_extras:
        for k in ('acc', 'distance_mean', 'distance_median', 'distance_percentile_5', 'distance_percentile_95'):
          rd[k] = torch.empty(n, device=dev)
      hd = {}
      if want_history:
        hd = {'sdist': torch.empty(n, S + 1, device=dev), 'weights': torch.empty(n, S, device=dev),
              'density': torch.empty(n, S, device=dev)}
        if l == L - 1:
          hd['rgb'] = torch.empty(n, S, 3, device=dev)
      o = outs[l]
      o.rgb = rd['rgb'].data_ptr()
      if compute_extras:
        o.acc, o.distance_mean = rd['acc'].data_ptr(), rd['distance_mean'].data_ptr()
        o.distance_median = rd['distance_median'].data_ptr()
        o.distance_p5, o.distance_p95 = rd['distance_percentile_5'].data_ptr(), rd['distance_percentile_95'].data_ptr()
      if want_history:
        o.sdist, o.weights, o.density = hd['sdist'].data_ptr(), hd['weights'].data_ptr(), hd['density'].data_ptr()
        if l == L - 1:
          o.rgbs = hd['rgb'].data_ptr()
      res.append(rd); hist.append(hd)
    jit = None if jitter is None else _f32(jitter, dev).reshape(L, n)
    with torch.cuda.device(dev):
      check(lib.hugs_forward(self._h, _ptr(params), C.byref(r), n, float(train_frac), _ptr(jit),
                             int(compute_extras), int(zero_glo), outs, self._stream()))
    self._keep = keep + [jit]
    return res, hist

  def render_frame(self, params: torch.Tensor, camera_set: '_lib.CameraSet', cam: int, width: int, height: int,
                   row0: int, row1: int, train_frac: float, zero_glo: bool = True, compute_extras: bool = True,
                   want_u8: bool = False, want_sse: bool = False) -> Dict[str, torch.Tensor]:
    """models.render_image for rows [row0, row1) of one camera of a device-resident dataset: ONE library call generates
    the rays and renders every chunk into frame-sized tensors (hugs_render_frame)."""
    dev, rows = self.device, row1 - row0
    out = {'rgb': torch.empty(rows, width, 3, device=dev), 'acc': torch.empty(rows, width, device=dev)}
    if compute_extras:
      out['distance_mean'] = torch.empty(rows, width, device=dev)
      out['distance_median'] = torch.empty(rows, width, device=dev)
    if want_u8:
      out['rgb_u8'] = torch.empty(rows, width, 3, device=dev, dtype=torch.uint8)
    if want_sse:
      out['sse'] = torch.zeros(2, device=dev, dtype=torch.float64)
    if rows == 0:
      return out
    fo = _lib.FrameOut()
    for k, v in out.items():
      setattr(fo, k, v.data_ptr())
    with torch.cuda.device(dev):
      check(lib.hugs_render_frame(self._h, _ptr(params), C.byref(camera_set), int(cam), int(width), int(height), int(row0),
                                  int(row1), float(train_frac), int(zero_glo), C.byref(fo), self._stream()))
    return out

  def loss_and_grad(self, params, rays, rgb_gt, train_frac, jitter, loss_cfg: '_lib.LossCfg',
                    grad_out: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None):
    r, keep, n = self._rays(rays)
    dev = self.device
    gt = _f32(rgb_gt, dev).reshape(n, 3)
    grad = torch.empty(self.n_params, device=dev) if grad_out is None else grad_out
    stats = torch.empty(16, device=dev) if stats_out is None else stats_out
    jit = None if jitter is None else _f32(jitter, dev).reshape(self.cfg.num_levels, n)
    with torch.cuda.device(dev):
      check(lib.hugs_loss_and_grad(self._h, _ptr(params), C.byref(r), _ptr(gt), n, float(train_frac), _ptr(jit),
                                   C.byref(loss_cfg), _ptr(grad), _ptr(stats), self._stream()))
    self._keep = keep + [gt, jit]
    return grad, stats

  def set_train_rng(self, seed: int, counter: int):
    """In-kernel jitter draws for loss_and_grad(jitter=None): seed 0 switches them off (deterministic sampling)."""
    check(lib.hugs_set_train_rng(self._h, int(seed) & (2 ** 64 - 1), int(counter) & (2 ** 64 - 1)))

  def set_grad_ready_event(self, event: Optional[torch.cuda.Event]):
    """`event` is recorded inside loss_and_grad once the NerfMLP_0 / GloEmbed_0 gradients are final."""
    self._grad_event = event            # keep it alive
    check(lib.hugs_set_grad_ready_event(self._h, None if event is None else C.c_void_p(event.cuda_event)))

  def adam_step(self, params, grad, mu, nu, adam_cfg: '_lib.AdamCfg', norms_out: Optional[torch.Tensor] = None,
                tensor_stats_out: Optional[torch.Tensor] = None):
    """clip + nan_to_num + Adam; `tensor_stats_out` ([len(layout), 5] fp32) also receives, per parameter tensor,
    {sum w^2, sum g^2, max |g|, sum delta^2, max |delta|} (hugs_adam_step_stats)."""
    with torch.cuda.device(self.device):
      if tensor_stats_out is None:
        check(lib.hugs_adam_step(self._h, _ptr(params), _ptr(grad), _ptr(mu), _ptr(nu), C.byref(adam_cfg),
                                 _ptr(norms_out), self._stream()))
      else:
        check(lib.hugs_adam_step_stats(self._h, _ptr(params), _ptr(grad), _ptr(mu), _ptr(nu), C.byref(adam_cfg),
                                       _ptr(norms_out), _ptr(tensor_stats_out), self._stream()))

  # ---- one field on caller-provided fenceposts (the torch twins, nerfacto/models/nerf.py:299-318) ------------
  def field_forward(self, params, rays, tdist, training: bool, zero_glo: bool = False, raw_out=None):
    """MLP.forward on the midpoints of `tdist` [n, S+1]: raw [n, S, 4] (pre-activation density, rgb)."""
    r, keep, n = self._rays(rays)
    tdist = _f32(tdist, self.device)
    S = tdist.shape[-1] - 1
    raw = torch.empty(n, S, 4, device=self.device) if raw_out is None else raw_out
    with torch.cuda.device(self.device):
      check(lib.hugs_field_forward(self._h, _ptr(params), C.byref(r), _ptr(tdist), n, S, int(training), int(zero_glo),
                                   _ptr(raw), self._stream()))
    self._keep = keep + [tdist]
    return raw

  def field_backward(self, params, rays, n_samples: int, d_raw, grad_out=None):
    r, keep, n = self._rays(rays)
    d_raw = _f32(d_raw, self.device)
    grad = torch.empty(self.n_params, device=self.device) if grad_out is None else grad_out
    with torch.cuda.device(self.device):
      check(lib.hugs_field_backward(self._h, _ptr(params), C.byref(r), n, int(n_samples), _ptr(d_raw), _ptr(grad),
                                    self._stream()))
    self._keep = keep + [d_raw]
    return grad

  # ---- measurement hooks ----------------------------------------------------------------------
  def profile(self, enable: bool):
    check(lib.hugs_profile_enable(self._h, int(enable)))

  def profile_read(self):
    ms = (C.c_float * len(_lib.KERNEL_CLASSES))()
    cnt = (C.c_int32 * len(_lib.KERNEL_CLASSES))()
    check(lib.hugs_profile_read(self._h, ms, cnt))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}

  @staticmethod
  def launch_count() -> int:
    return int(lib.hugs_launch_count())

  # ---- operator-level calls (parity tests) ---------------------------------------------------
  def sample_intervals(self, t, w_logits, u_base, jitter, max_jitter, n_samples, domain, want_idx=False):
    dev = self.device
    t, w_logits, u_base = _f32(t, dev), _f32(w_logits, dev), _f32(u_base, dev)
    n, nb = w_logits.shape
    out = torch.empty(n, n_samples + 1, device=dev)
    idx = torch.empty(n, n_samples, device=dev, dtype=torch.int32) if want_idx else None
    jit = None if jitter is None else _f32(jitter, dev).reshape(n)
    check(lib.hugs_sample_intervals(_ptr(t), _ptr(w_logits), _ptr(u_base), _ptr(jit), float(max_jitter), n, nb,
                                    n_samples, float(domain[0]), float(domain[1]), _ptr(out), _ptr(idx),
                                    self._stream()))
    torch.cuda.synchronize(dev)
    return (out, idx) if want_idx else out

  def invert_cdf(self, t, cw, u):
    dev = self.device
    t, cw, u = _f32(t, dev), _f32(cw, dev), _f32(u, dev)
    n, nb1 = t.shape
    ns = u.shape[-1]
    out = torch.empty(n, ns, device=dev)
    idx = torch.empty(n, ns, device=dev, dtype=torch.int32)
    check(lib.hugs_invert_cdf(_ptr(t), _ptr(cw), _ptr(u), n, nb1 - 1, ns, _ptr(out), _ptr(idx), self._stream()))
    torch.cuda.synchronize(dev)
    return out, idx

  def max_dilate_weights(self, t, w, dilation, domain):
    dev = self.device
    t, w = _f32(t, dev), _f32(w, dev)
    n, nb = w.shape
    to = torch.empty(n, 3 * nb - 1, device=dev)
    wo = torch.empty(n, 3 * nb - 2, device=dev)
    check(lib.hugs_max_dilate_weights(_ptr(t), _ptr(w), n, nb, float(dilation), float(domain[0]), float(domain[1]),
                                      _ptr(to), _ptr(wo), self._stream()))
    torch.cuda.synchronize(dev)
    return to, wo

  def alpha_composite(self, raw_density, raw_rgb, tdist, directions, far, compute_extras=True):
    dev = self.device
    raw_density, tdist, directions = _f32(raw_density, dev), _f32(tdist, dev), _f32(directions, dev)
    raw_rgb = None if raw_rgb is None else _f32(raw_rgb, dev)
    far = _f32(far, dev).reshape(-1)
    n, S = raw_density.shape
    o = _lib.LevelOut()
    out = {'rgb': torch.empty(n, 3, device=dev), 'weights': torch.empty(n, S, device=dev)}
    for k in ('acc', 'distance_mean', 'distance_median', 'distance_percentile_5', 'distance_percentile_95'):
      out[k] = torch.empty(n, device=dev)
    o.rgb, o.weights, o.acc = out['rgb'].data_ptr(), out['weights'].data_ptr(), out['acc'].data_ptr()
    o.distance_mean, o.distance_median = out['distance_mean'].data_ptr(), out['distance_median'].data_ptr()
    o.distance_p5, o.distance_p95 = out['distance_percentile_5'].data_ptr(), out['distance_percentile_95'].data_ptr()
    check(lib.hugs_alpha_composite(self._h, _ptr(raw_density), _ptr(raw_rgb), _ptr(tdist), None, _ptr(directions),
                                   _ptr(far), n, S, int(compute_extras), C.byref(o), self._stream()))
    torch.cuda.synchronize(dev)
    return out

  def debug_encode_bf16(self, rays, tdist, contract: bool):
    """The throughput-mode bf16 IPE encoder on its own: [n*S, 512] bf16, engine column order (test hook)."""
    r, keep, n = self._rays(rays)
    tdist = _f32(tdist, self.device)
    S = tdist.shape[-1] - 1
    out = torch.empty(n * S, 512, device=self.device, dtype=torch.bfloat16)
    check(lib.hugs_debug_encode_bf16(self._h, C.byref(r), _ptr(tdist), n, S, int(contract), _ptr(out), self._stream()))
    torch.cuda.synchronize(self.device)
    return out

  def ipe_features(self, rays, tdist, contract: bool):
    r, keep, n = self._rays(rays)
    tdist = _f32(tdist, self.device)
    S = tdist.shape[-1] - 1
    ndeg = self.cfg.max_deg_point - self.cfg.min_deg_point
    fd = 3 + 6 * ndeg if self.cfg.encoding == 'point_pe' else 2 * self.num_basis * ndeg
    out = torch.empty(n * S, fd, device=self.device)
    check(lib.hugs_ipe_features(self._h, C.byref(r), _ptr(tdist), n, S, int(contract), _ptr(out), self._stream()))
    torch.cuda.synchronize(self.device)
    return out.reshape(n, S, fd)
