"""Operators of the torch twins as torch.autograd Functions over the C ABI (include/hugs_b200.h, hugs_nf_* / hugs_field_*).

PyTorch holds the device memory and records the autograd graph; every number is computed by libhugs_b200.so.
Reference semantics: /root/reference/nerfacto/utils/ray_utils.py, models/nerf.py (file:line in the header).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import lib, check
from ..engine import Engine, EngineConfig

SPACING = {'uniform': 0, 'reciprocal': 1, 'piecewise': 3}          # hugs_raydist_fn
DENSITY_ACT = {'softplus': 0, 'trunc_exp': 1, 'relu': 2}            # hugs_density_act
LOSS_TYPE = {'charb': 0, 'mse': 1}                                  # hugs_data_loss


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
  return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
  return t.detach().to(torch.float32).contiguous()


def _require_cuda(t: torch.Tensor, what: str):
  if not t.is_cuda:
    raise RuntimeError(f'{what}: tensors must live on a CUDA device (nerf_hugs_b200 has no CPU path)')


_U_CACHE: Dict[Tuple, Tuple[torch.Tensor, float]] = {}


def sample_u(num_samples: int, perturb: bool, device) -> Tuple[torch.Tensor, float]:
  """The `u` grid of ray_utils.sample (ray_utils.py:147-159): (u_base [ns] on `device`, max_jitter)."""
  key = (num_samples, bool(perturb), str(device))
  if key not in _U_CACHE:
    eps = float(torch.finfo(torch.float32).eps)
    if perturb:
      u_max = eps + (1. - eps) / num_samples
      max_jitter = (1. - u_max) / (num_samples - 1) - eps
      u = torch.linspace(0, 1. - u_max, num_samples, dtype=torch.float32)
    else:
      pad = 1 / (2 * num_samples)
      max_jitter = 0.
      u = torch.linspace(pad, 1. - pad - eps, num_samples, dtype=torch.float32)
    _U_CACHE[key] = (u.to(device), float(np.float32(max_jitter)))
  return _U_CACHE[key]


@torch.no_grad()
def sample_intervals(spacing_bins, weights, anneal, padding, num_samples, perturb, single_jitter, domain,
                     spacing_fn='uniform', near=None, far=None, jitter=None):
  """ray_utils.sample_intervals (ray_utils.py:196-223) (+ s_to_t of the new fenceposts when near / far are given).

  jitter: optional caller-provided uniform draws ([n, 1] when single_jitter else [n, num_samples]) in place of torch.rand
  (ray_utils.py:151-152), so that a run can be repeated on given draws.  Returns (new_bins, euclidean_bins or None).
  """
  _require_cuda(spacing_bins, 'sample_intervals')
  dev = spacing_bins.device
  bins, w = _f32c(spacing_bins), _f32c(weights)
  n, nb = w.shape
  u_base, max_jitter = sample_u(num_samples, perturb, dev)
  stride = 1
  if perturb:
    d = 1 if single_jitter else num_samples
    if jitter is None:
      jitter = torch.rand((n, d), dtype=torch.float32, device=dev)
    jitter = _f32c(jitter).reshape(n, d)
    stride = d
  else:
    jitter = None
  out = torch.empty(n, num_samples + 1, device=dev)
  t_out = torch.empty_like(out) if near is not None else None
  near_c = None if near is None else _f32c(near).reshape(n)
  far_c = None if far is None else _f32c(far).reshape(n)
  with torch.cuda.device(dev):
    check(lib.hugs_nf_sample_intervals(_ptr(bins), _ptr(w), _ptr(u_base), _ptr(jitter), stride, max_jitter,
                                       float(anneal), float(padding), n, nb, num_samples, float(domain[0]),
                                       float(domain[1]), SPACING[spacing_fn], _ptr(near_c), _ptr(far_c), _ptr(out),
                                       _ptr(t_out), _stream(dev)))
  return out, t_out


@torch.no_grad()
def merge_bins(bins_a, bins_b, domain, spacing_fn='uniform', near=None, far=None):
  """nerf.py:287-295: fenceposts around the sorted union of the centres of two fencepost sets (+ s_to_t)."""
  _require_cuda(bins_a, 'merge_bins')
  dev = bins_a.device
  a, b = _f32c(bins_a), _f32c(bins_b)
  n, na, nb = a.shape[0], a.shape[1] - 1, b.shape[1] - 1
  out = torch.empty(n, na + nb + 1, device=dev)
  t_out = torch.empty_like(out) if near is not None else None
  near_c = None if near is None else _f32c(near).reshape(n)
  far_c = None if far is None else _f32c(far).reshape(n)
  with torch.cuda.device(dev):
    check(lib.hugs_nf_merge_bins(_ptr(a), na, _ptr(b), nb, n, float(domain[0]), float(domain[1]), SPACING[spacing_fn],
                                 _ptr(near_c), _ptr(far_c), _ptr(out), _ptr(t_out), _stream(dev)))
  return out, t_out


def render_cfg(opaque_background, density_activation, density_bias, rgb_premultiplier=1., rgb_bias=0., rgb_padding=0.):
  c = _lib.NfRenderCfg()
  c.opaque_background = int(bool(opaque_background))
  c.density_activation = DENSITY_ACT[density_activation]
  c.density_bias, c.rgb_premultiplier, c.rgb_bias, c.rgb_padding = density_bias, rgb_premultiplier, rgb_bias, rgb_padding
  return c


def composite_forward(cfg, raw, tdist, directions, bg_rgb):
  """density_to_weight + render_features + render_depth (ray_utils.py:226-249,295-312,336-346) of raw [n, S, C]."""
  dev = raw.device
  n, S, Cc = raw.shape
  weights = torch.empty(n, S, device=dev)
  rgb = torch.empty(n, 3, device=dev) if Cc == 4 else None
  depth = torch.empty(n, device=dev)
  acc = torch.empty(n, device=dev)
  steps_max = torch.full((1,), float('-inf'), device=dev)
  with torch.cuda.device(dev):
    st = _stream(dev)
    check(lib.hugs_nf_composite(C.byref(cfg), _ptr(raw), Cc, _ptr(tdist), _ptr(directions), _ptr(bg_rgb), n, S,
                                _ptr(weights), _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(steps_max), st))
    check(lib.hugs_nf_clip_depth(_ptr(depth), _ptr(steps_max), n, st))      # quirk B7: batch-wide maximum
  return weights, rgb, depth, acc, steps_max


def composite_backward(cfg, raw, tdist, directions, bg_rgb, steps_max, d_weights, d_rgb, d_depth, d_acc):
  dev = raw.device
  n, S, Cc = raw.shape
  d_raw = torch.empty_like(raw)
  gs = [None if g is None else _f32c(g) for g in (d_weights, d_rgb, d_depth, d_acc)]
  with torch.cuda.device(dev):
    check(lib.hugs_nf_composite_bwd(C.byref(cfg), _ptr(raw), Cc, _ptr(tdist), _ptr(directions), _ptr(bg_rgb), n, S,
                                    _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), _ptr(steps_max), _ptr(d_raw),
                                    _stream(dev)))
  return d_raw


class FieldEngine:
  """One field MLP (nerf.py:632-860) on the tensor-core engine: a num_levels = 1 handle with a point encoding, plus the
  table-driven copies between the torch parameters ([out, in] nn.Linear weights) and the handle's flat flax-layout buffer."""

  def __init__(self, ecfg: EngineConfig, device, linears: Sequence[torch.nn.Linear],
               embedding: Optional[torch.nn.Embedding]):
    self.ecfg, self.device = ecfg, torch.device(device)
    self.engine = Engine(ecfg, np.zeros((3, 1), np.float32), device=self.device)
    self.params: List[torch.nn.Parameter] = []
    self.spec: List[Tuple[int, int, int, int]] = []      # (flat_off, rows, cols, transpose) per entry of self.params
    lay = {name: (off, rows, cols) for name, off, rows, cols, _ in self.engine.layout}
    for i, lin in enumerate(linears):
      off, rows, cols = lay[f'NerfMLP_0/Dense_{i}/kernel']
      assert tuple(lin.weight.shape) == (cols, rows), (i, tuple(lin.weight.shape), rows, cols)
      self.params.append(lin.weight); self.spec.append((off, rows, cols, 1))
      off, rows, cols = lay[f'NerfMLP_0/Dense_{i}/bias']
      self.params.append(lin.bias); self.spec.append((off, rows, cols, 0))
    if embedding is not None:
      off, rows, cols = lay['GloEmbed_0/embedding']
      assert tuple(embedding.weight.shape) == (rows, cols)
      self.params.append(embedding.weight); self.spec.append((off, rows, cols, 0))
    self.n_linear_params = 2 * len(linears)
    self.flat = torch.zeros(self.engine.n_params, device=self.device)
    self.gflat = torch.zeros(self.engine.n_params, device=self.device)
    self._table_key = None
    self._table = None
    self._version_key = None
    self._gtable_key = None
    self._gtable = None

  def _make_table(self, ptrs):
    """one hugs_tensor_copy per parameter; `ptrs` are addresses (import) or byte offsets into a gradient buffer (export)"""
    arr = (_lib.TensorCopy * len(ptrs))()
    for e, ptr, (off, rows, cols, tr) in zip(arr, ptrs, self.spec):
      e.ptr, e.flat_off, e.rows, e.cols, e.transpose = ptr, off, rows, cols, tr
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(self.device)

  def sync_params(self, tensors: Sequence[torch.Tensor], force: bool = False):
    """tensors (same order as self.params) -> flat buffer -> packed operands.  Skipped only when every tensor IS the
    long-lived parameter with an unchanged version counter (a temporary may reuse the address of an earlier one, and
    `p.data` edits do not bump the counter: training steps and stand-in tensors always refresh)."""
    for t in tensors:
      if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError('field parameters must be contiguous fp32 CUDA tensors (model.to(device) first)')
    key = tuple(t.data_ptr() for t in tensors)
    if key != self._table_key:
      self._table, self._table_key, self._version_key = self._make_table(key), key, None
    vkey = tuple(t._version for t in tensors)
    if force or any(t is not p for t, p in zip(tensors, self.params)):
      self._version_key = None
    if vkey != self._version_key:
      with torch.cuda.device(self.device):
        check(lib.hugs_params_copy(_ptr(self._table), len(tensors), _ptr(self.flat), 0, None, _stream(self.device)))
      self.engine.params_changed(self.flat)
      self._version_key = vkey

  def export_grads(self) -> List[torch.Tensor]:
    """flat flax-layout gradient -> one tensor per parameter in torch layout (views of one buffer)."""
    sizes = [rows * cols for _, rows, cols, _ in self.spec]
    buf = torch.empty(sum(sizes), device=self.device)       # a fresh buffer per backward pass: autograd may keep it as .grad
    outs, o = [], 0
    for p, sz in zip(self.params, sizes):
      outs.append(buf[o:o + sz].view(p.shape)); o += sz
    if self._gtable is None:                                # constant: byte offsets into `buf` (no per-step host -> device copy)
      self._gtable = self._make_table([4 * int(sum(sizes[:i])) for i in range(len(sizes))])
    with torch.cuda.device(self.device):
      check(lib.hugs_params_copy(_ptr(self._gtable), len(outs), _ptr(self.gflat), 1, _ptr(buf), _stream(self.device)))
    return outs


class _RenderField(torch.autograd.Function):
  """One field of Model.forward_rays (nerf.py:298-331): field MLP on the interval midpoints -> density_to_weight ->
  render_features / render_depth.  Inputs after `n_fixed` are the field's parameters (autograd leaves)."""

  @staticmethod
  def forward(ctx, field: FieldEngine, rays: Dict[str, torch.Tensor], tdist, bg_rgb, cfg, training: bool, zero_glo: bool,
              *params):
    field.sync_params(params, force=training)
    raw = field.engine.field_forward(field.flat, rays, tdist, training=training, zero_glo=zero_glo)
    dirs = _f32c(rays['directions'])
    bg = None if bg_rgb is None else _f32c(bg_rgb)
    weights, rgb, depth, acc, steps_max = composite_forward(cfg, raw, tdist, dirs, bg)
    ctx.field, ctx.rays, ctx.cfg, ctx.n_params = field, rays, cfg, len(params)
    ctx.save_for_backward(raw, tdist, dirs, bg, steps_max)
    ctx.set_materialize_grads(False)
    return rgb, depth, acc, weights

  @staticmethod
  def backward(ctx, d_rgb, d_depth, d_acc, d_weights):
    raw, tdist, dirs, bg, steps_max = ctx.saved_tensors
    field = ctx.field
    d_raw = composite_backward(ctx.cfg, raw, tdist, dirs, bg, steps_max, d_weights, d_rgb, d_depth, d_acc)
    field.engine.field_backward(field.flat, ctx.rays, raw.shape[1], d_raw, grad_out=field.gflat)
    grads = field.export_grads()
    return (None,) * 7 + tuple(grads)


def render_field(field: FieldEngine, rays, tdist, bg_rgb, cfg, training: bool, zero_glo: bool = False, params=None):
  params = field.params if params is None else params
  return _RenderField.apply(field, rays, tdist, bg_rgb, cfg, training, zero_glo, *params)


class _RgbLoss(torch.autograd.Function):
  """rgb_loss_mult * sum(lossmult * rgb_loss((pred - gt)^2)) / max(sum lossmult, eps) (nerf.py:404-461).  Returns the loss and
  the equally normalised squared error (`mse`, detached)."""

  @staticmethod
  def forward(ctx, pred, gt, static_mask, transient_weight, loss_type, padding, mult):
    dev = pred.device
    p, g = _f32c(pred).reshape(-1, 3), _f32c(gt).reshape(-1, 3)
    n = p.shape[0]
    m = None if static_mask is None else _f32c(static_mask).reshape(n)
    sums = torch.empty(3, device=dev)
    dl = torch.empty(n, 3, device=dev)
    with torch.cuda.device(dev):
      check(lib.hugs_nf_rgb_loss(_ptr(p), _ptr(g), _ptr(m), float(transient_weight), LOSS_TYPE[loss_type], float(padding),
                                 n, _ptr(sums), _ptr(dl), _stream(dev)))
    ctx.save_for_backward(dl, sums)
    ctx.mult, ctx.shape = float(mult), pred.shape
    denom = sums[2].clamp_min(torch.finfo(torch.float32).eps)
    out = torch.stack([mult * sums[0] / denom, sums[1] / denom])
    return out

  @staticmethod
  def backward(ctx, d_out):
    dl, sums = ctx.saved_tensors
    dev = dl.device
    up = _f32c(d_out)[:1].contiguous()
    d_pred = torch.empty_like(dl)
    with torch.cuda.device(dev):
      check(lib.hugs_nf_rgb_loss_bwd(_ptr(dl), _ptr(sums), _ptr(up), ctx.mult, dl.shape[0], _ptr(d_pred), _stream(dev)))
    return d_pred.view(ctx.shape), None, None, None, None, None, None


def rgb_loss(pred, gt, static_mask, transient_weight, loss_type, padding, mult):
  """-> (loss, mse): loss carries the gradient, mse is detached."""
  out = _RgbLoss.apply(pred, gt, static_mask, transient_weight, loss_type, padding, mult)
  return out[0], out[1].detach()


# ------------------------------------------------------------------------------------------------ hash-grid fields
class HashFieldEngine:
  """One hash-grid field (nerfacto.py:643-1008) behind hugs_hashfield_*: the handle, the table-driven copies between the
  torch parameters and the flat [in, out] MLP buffer.  The hash table itself (`params` of the encoder, tcnn layout) is used
  in place, never copied."""

  def __init__(self, desc: '_lib.HashFieldDesc', device, grid_param: torch.nn.Parameter,
               entries: Sequence[Tuple[str, torch.Tensor, int, int, int]]):
    """entries: (flat tensor name, torch tensor, element offset into it, transpose, ld) in any order."""
    if not torch.cuda.is_available():
      raise RuntimeError('nerf_hugs_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    self.device = torch.device(device)
    self.desc = desc
    self._h = C.c_void_p()
    with torch.cuda.device(self.device):
      check(lib.hugs_hashfield_create(C.byref(desc), C.byref(self._h)))
    self.grid_floats = int(lib.hugs_hashfield_grid_floats(self._h))
    self.mlp_floats = int(lib.hugs_hashfield_mlp_floats(self._h))
    cnt = C.c_int32()
    check(lib.hugs_hashfield_layout(self._h, None, 0, C.byref(cnt)))
    arr = (_lib.TensorDesc * cnt.value)()
    check(lib.hugs_hashfield_layout(self._h, arr, cnt.value, C.byref(cnt)))
    self.layout = {t.name.decode(): (int(t.offset), int(t.rows), int(t.cols)) for t in arr}
    assert grid_param.numel() == self.grid_floats, (grid_param.numel(), self.grid_floats)
    self.grid_param = grid_param
    self.entries = list(entries)
    self.is_density_only = desc.geo_feat_dim == 0
    self.flat = torch.zeros(self.mlp_floats, device=self.device)
    self.gflat = torch.zeros(self.mlp_floats, device=self.device)
    self._table_key = self._table = self._version_key = self._gtable = None

  def close(self):
    if self._h:
      lib.hugs_hashfield_destroy(self._h)
      self._h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def level_info(self, level):
    s, r, o, e = C.c_float(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib.hugs_hashfield_level_info(self._h, level, C.byref(s), C.byref(r), C.byref(o), C.byref(e)))
    return float(s.value), int(r.value), int(o.value), int(e.value)

  def _make_table(self, ptrs):
    """one hugs_tensor_copy per entry; ptrs[i] = address (or byte offset into a gradient buffer) of entry i's torch tensor"""
    arr = (_lib.TensorCopy * len(self.entries))()
    for e, (name, _, off, tr, ld), ptr in zip(arr, self.entries, ptrs):
      foff, rows, cols = self.layout[name.split('#')[0]]
      if '#' in name:      # "tensor#row0:rows": a block of rows of the flat tensor
        r0, nr = (int(v) for v in name.split('#')[1].split(':'))
        foff, rows = foff + r0 * cols, nr
      e.ptr, e.flat_off, e.rows, e.cols, e.transpose, e.ld = ptr + 4 * off, foff, rows, cols, tr, ld
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)

  def sync_params(self, force: bool = False, override: Optional[Dict[int, torch.Tensor]] = None):
    tensors = [e[1] for e in self.entries]
    if override:
      tensors = [override.get(id(t), t) for t in tensors]
    for t in tensors + [self.grid_param]:
      if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError('field parameters must be contiguous fp32 CUDA tensors (model.to(device) first)')
    key = tuple(t.data_ptr() for t in tensors)
    if key != self._table_key:
      self._table, self._table_key, self._version_key = self._make_table(key), key, None
    vkey = tuple(t._version for t in tensors)
    if force or override:
      self._version_key = None
    if vkey != self._version_key:
      with torch.cuda.device(self.device):
        check(lib.hugs_params_copy(_ptr(self._table), len(tensors), _ptr(self.flat), 0, None, _stream(self.device)))
        check(lib.hugs_hashfield_params_changed(self._h, _ptr(self.flat), _stream(self.device)))
      self._version_key = vkey

  def _rays(self, rays):
    r = _lib.Rays()
    keep = []
    n = rays['origins'].shape[0]
    for k in ('origins', 'directions', 'viewdirs'):
      if rays.get(k) is not None:
        t = _f32c(rays[k]).reshape(n, 3)
        keep.append(t); setattr(r, k, t.data_ptr())
    if rays.get('embed_idx') is not None:
      t = rays['embed_idx'].detach().to(torch.int32).contiguous().reshape(n)
      keep.append(t); r.embed_idx = t.data_ptr()
    return r, keep, n

  def encode(self, rays, tdist):
    r, keep, n = self._rays(rays)
    tdist = _f32c(tdist)
    S = tdist.shape[1] - 1
    out = torch.empty(n * S, 2 * self.desc.n_levels, device=self.device)
    with torch.cuda.device(self.device):
      check(lib.hugs_hashfield_encode(self._h, _ptr(self.grid_param.detach()), C.byref(r), _ptr(tdist), n, S, _ptr(out),
                                      _stream(self.device)))
    return out

  def forward(self, rays, tdist, training: bool, zero_app: bool = False):
    r, keep, n = self._rays(rays)
    S = tdist.shape[1] - 1
    raw = torch.empty((n, S, 1 if self.is_density_only else 4), device=self.device)
    with torch.cuda.device(self.device):
      check(lib.hugs_hashfield_forward(self._h, _ptr(self.grid_param.detach()), _ptr(self.flat), C.byref(r), _ptr(tdist), n, S,
                                       int(training), int(zero_app), _ptr(raw), _stream(self.device)))
    self._keep = keep
    return raw

  def backward(self, rays, tdist, d_raw):
    r, keep, n = self._rays(rays)
    S = tdist.shape[1] - 1
    grid_grad = torch.zeros(self.grid_floats, device=self.device)
    with torch.cuda.device(self.device):
      check(lib.hugs_hashfield_backward(self._h, _ptr(self.grid_param.detach()), _ptr(self.flat), C.byref(r), _ptr(tdist), n, S,
                                        _ptr(d_raw), _ptr(grid_grad), _ptr(self.gflat), _stream(self.device)))
    self._keep = keep
    return grid_grad

  def export_grads(self, shapes: Sequence[torch.Size]) -> List[torch.Tensor]:
    """flat MLP gradient -> one tensor per DISTINCT torch parameter of self.entries (several entries may address blocks of
    one parameter)."""
    distinct, seen = [], {}
    for _, t, _, _, _ in self.entries:
      if id(t) not in seen:
        seen[id(t)] = len(distinct); distinct.append(t)
    sizes = [t.numel() for t in distinct]
    buf = torch.empty(sum(sizes), device=self.device)       # a fresh buffer per backward pass (the entries cover every element)
    outs, o = [], 0
    for t, sz in zip(distinct, sizes):
      outs.append(buf[o:o + sz].view(t.shape)); o += sz
    if self._gtable is None:                                # constant: byte offsets into `buf`
      starts = [4 * int(sum(sizes[:i])) for i in range(len(sizes))]
      self._gtable = self._make_table([starts[seen[id(t)]] for _, t, _, _, _ in self.entries])
    with torch.cuda.device(self.device):
      check(lib.hugs_params_copy(_ptr(self._gtable), len(self.entries), _ptr(self.gflat), 1, _ptr(buf), _stream(self.device)))
    return distinct, outs


class _RenderHashField(torch.autograd.Function):
  """One level of nerfacto's Model.forward_rays (nerfacto.py:329-371): hash-grid field -> density_to_weight ->
  render_features / render_depth.  Differentiable inputs: the hash table and the field's MLP parameters."""

  @staticmethod
  def forward(ctx, field: HashFieldEngine, rays, tdist, bg_rgb, cfg, training: bool, zero_app: bool, override, grid, *params):
    field.sync_params(force=training, override=override)
    raw = field.forward(rays, tdist, training, zero_app)
    dirs = _f32c(rays['directions'])
    bg = None if (bg_rgb is None or field.is_density_only) else _f32c(bg_rgb)
    weights, rgb, depth, acc, steps_max = composite_forward(cfg, raw, tdist, dirs, bg)
    ctx.field, ctx.rays, ctx.cfg = field, rays, cfg
    ctx.param_ids = [id(p) for p in params]
    ctx.save_for_backward(raw, tdist, dirs, bg, steps_max)
    ctx.set_materialize_grads(False)
    if rgb is None:
      return depth, acc, weights
    return rgb, depth, acc, weights

  @staticmethod
  def backward(ctx, *grads):
    raw, tdist, dirs, bg, steps_max = ctx.saved_tensors
    field = ctx.field
    if raw.shape[2] == 1:
      d_depth, d_acc, d_weights = grads
      d_rgb = None
    else:
      d_rgb, d_depth, d_acc, d_weights = grads
    d_raw = composite_backward(ctx.cfg, raw, tdist, dirs, bg, steps_max, d_weights, d_rgb, d_depth, d_acc)
    grid_grad = field.backward(ctx.rays, tdist, d_raw)
    distinct, outs = field.export_grads(None)
    by_id = {id(t): g for t, g in zip(distinct, outs)}
    return (None,) * 8 + (grid_grad,) + tuple(by_id.get(i) for i in ctx.param_ids)


def render_hash_field(field: HashFieldEngine, rays, tdist, bg_rgb, cfg, training, zero_app, grid, params, override=None):
  return _RenderHashField.apply(field, rays, tdist, bg_rgb, cfg, training, zero_app, override, grid, *params)


class _ScaledLoss(torch.autograd.Function):
  """value = sum / count of a per-ray loss kernel; backward = upstream / count * stored gradient (hugs_nf_scale)."""

  @staticmethod
  def forward(ctx, w_for_grad, total, grad, count):
    ctx.save_for_backward(grad)
    ctx.count = float(count)
    return total[0] / count

  @staticmethod
  def backward(ctx, up):
    (grad,) = ctx.saved_tensors
    dev = grad.device
    out = torch.empty_like(grad)
    upc = _f32c(up).reshape(1)
    with torch.cuda.device(dev):
      check(lib.hugs_nf_scale(_ptr(grad), _ptr(upc), 1.0 / ctx.count, grad.numel(), _ptr(out), _stream(dev)))
    return out, None, None, None


def distortion_loss(weights_list, spacing_bins_list):
  """loss_utils.distortion_loss (loss_utils.py:80-84): mean over rays of lossfun_distortion on the final level."""
  c, w = _f32c(spacing_bins_list[-1]), weights_list[-1]
  wc = _f32c(w)
  n, S = wc.shape
  dev = wc.device
  total, grad = torch.empty(1, device=dev), torch.empty(n, S, device=dev)
  with torch.cuda.device(dev):
    check(lib.hugs_nf_distortion_loss(_ptr(c), _ptr(wc), n, S, _ptr(total), _ptr(grad), _stream(dev)))
  return _ScaledLoss.apply(w, total, grad, n)


def interlevel_loss(weights_list, spacing_bins_list):
  """loss_utils.interlevel_loss (loss_utils.py:48-63): the final level's (detached) histogram bounds every proposal one."""
  c, w = _f32c(spacing_bins_list[-1]), _f32c(weights_list[-1])
  n, S = w.shape
  dev = w.device
  loss = 0.0
  for cp, wp in zip(spacing_bins_list[:-1], weights_list[:-1]):
    cpc, wpc = _f32c(cp), _f32c(wp)
    Sp = wpc.shape[1]
    total, grad = torch.empty(1, device=dev), torch.empty(n, Sp, device=dev)
    with torch.cuda.device(dev):
      check(lib.hugs_nf_interlevel_loss(_ptr(c), _ptr(w), S, _ptr(cpc), _ptr(wpc), Sp, n, _ptr(total), _ptr(grad), _stream(dev)))
    loss = loss + _ScaledLoss.apply(wp, total, grad, n * S)
  return loss
