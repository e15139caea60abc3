"""models.Model / construct_model / render_image with the reference signatures
(MipNeRF360/internal/models.py:47-357, :568-649), executing on the hugs_b200 engine."""
import dataclasses
import math
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

from .. import engine as _engine
from . import configs as _configs
from . import geopoly
from . import utils


def _bindings_of(config) -> _configs.Bindings:
  b = getattr(config, 'bindings', None)
  if b is None:
    b = _configs.Bindings(config=config)
    config.bindings = b
  return b


def engine_config(config, max_rays: int) -> _engine.EngineConfig:
  """Maps the gin-bound Model / NerfMLP / PropMLP fields onto hugs_model_desc."""
  b = _bindings_of(config)
  m, nm, pm = b.model, b.nerf_mlp, b.prop_mlp
  if config.transient_type not in (None, 'withmask'):
    raise NotImplementedError(
        f"transient_type={config.transient_type!r}: only None and 'withmask' (HuGS static masks) are on this path; "
        'robustnerf / nerfw / hanerf are the competing baselines of the reference (SURVEY.md §2.1)')
  if m.bg_intensity_range[0] != m.bg_intensity_range[1]:
    raise NotImplementedError('random background colours (bg_intensity_range min != max) are not supported')
  if m.disable_integration or not m.use_viewdirs or not m.single_jitter or not m.stop_level_grad:
    raise NotImplementedError('disable_integration / use_viewdirs=False / single_jitter=False / stop_level_grad=False')
  if nm.disable_rgb or not pm.disable_rgb:
    raise NotImplementedError('NerfMLP must output rgb and PropMLP.disable_rgb must be True (all shipped gins)')
  for f in ('min_deg_point', 'max_deg_point', 'skip_layer', 'basis_shape', 'basis_subdivisions', 'density_bias'):
    if getattr(nm, f) != getattr(pm, f):
      raise NotImplementedError(f'NerfMLP.{f} != PropMLP.{f}')
  if nm.net_depth_viewdirs != 1:
    raise NotImplementedError('net_depth_viewdirs != 1')
  return _engine.EngineConfig(
      num_levels=m.num_levels, num_prop_samples=m.num_prop_samples, num_nerf_samples=m.num_nerf_samples,
      nerf_depth=nm.net_depth, nerf_width=nm.net_width, prop_depth=pm.net_depth, prop_width=pm.net_width,
      bottleneck_width=nm.bottleneck_width, view_width=nm.net_width_viewdirs, skip_layer=nm.skip_layer,
      min_deg_point=nm.min_deg_point, max_deg_point=nm.max_deg_point, deg_view=nm.deg_view,
      raydist_fn=m.raydist_fn, ray_shape=m.ray_shape, nerf_contract=nm.warp_fn == 'contract',
      prop_contract=pm.warp_fn == 'contract', opaque_background=m.opaque_background,
      bg_intensity=float(m.bg_intensity_range[0]), anneal_slope=m.anneal_slope,
      dilation_multiplier=m.dilation_multiplier, dilation_bias=m.dilation_bias,
      resample_padding=m.resample_padding, near_anneal_rate=m.near_anneal_rate,
      near_anneal_init=m.near_anneal_init, num_glo_features=m.num_glo_features, num_embeddings=m.num_embeddings,
      density_bias=nm.density_bias, rgb_premultiplier=nm.rgb_premultiplier, rgb_bias=nm.rgb_bias,
      rgb_padding=nm.rgb_padding, precision=config.precision, max_rays=max_rays)


class Model:
  """models.Model (models.py:47-330).  `apply` mirrors flax's `model.apply(variables, rng, rays, ...)`."""

  def __init__(self, config, max_rays: Optional[int] = None, device=None):
    self.config = config
    b = _bindings_of(config)
    self.bindings = b
    for f in dataclasses.fields(b.model):
      setattr(self, f.name, getattr(b.model, f.name))
    self.max_rays = int(max_rays or max(config.batch_size, config.render_chunk_size))
    basis = geopoly.generate_basis(b.nerf_mlp.basis_shape, b.nerf_mlp.basis_subdivisions).T   # pos_basis_t
    self.pos_basis_t = np.ascontiguousarray(basis, dtype=np.float32)
    self.engine = _engine.Engine(engine_config(config, self.max_rays), self.pos_basis_t, device=device)
    self._packed_version = None

  # -- parameters -------------------------------------------------------------------------------
  def init(self, rng: int):
    """flax Dense(kernel_init=he_uniform) + zero bias (models.py:432-433); nn.Embed default init."""
    g = torch.Generator().manual_seed(int(rng))
    tree: Dict[str, Any] = {}
    for name, _, rows, cols, _ in self.engine.layout:
      parts = name.split('/')
      node = tree
      for q in parts[:-1]:
        node = node.setdefault(q, {})
      if parts[-1] == 'kernel':
        bound = math.sqrt(6.0 / rows)
        node['kernel'] = (torch.rand(rows, cols, generator=g) * 2 - 1) * bound
      elif parts[-1] == 'bias':
        node['bias'] = torch.zeros(cols)
      else:
        node['embedding'] = torch.randn(rows, cols, generator=g) / math.sqrt(cols)
    return {'params': tree}

  def flat_params(self, variables) -> torch.Tensor:
    tree = variables['params'] if isinstance(variables, dict) and 'params' in variables else variables
    return self.engine.flatten_params(tree)

  def _ensure_packed(self, flat: torch.Tensor):
    """Re-derives the packed bf16 operand copies unless `flat` is the very tensor (same object, same version counter)
    they were derived from.  The tensor is held by a strong reference: a temporary built from a parameter tree can
    then never be freed and have its address handed to the next, different tree."""
    ref = self._packed_version
    if ref is None or ref[0] is not flat or ref[1] != flat._version:
      self.engine.params_changed(flat)
      self._packed_version = (flat, flat._version)

  # -- forward ----------------------------------------------------------------------------------
  def apply(self, variables, rng, rays, train_frac, compute_extras, zero_glo=False, zero_tra=False):
    """Model.__call__ (models.py:74-330): returns (renderings, ray_history).

    variables: flat fp32 device tensor (TrainState.params) or a flax-style tree.
    rng: None (deterministic) or a torch.Generator on the engine device (one uniform draw per level & ray).
    """
    del zero_tra
    flat = variables if torch.is_tensor(variables) else self.flat_params(variables)
    self._ensure_packed(flat)
    rd = rays.as_dict() if isinstance(rays, utils.Rays) else dict(rays)
    lead = tuple(rd['origins'].shape[:-1])
    n = int(np.prod(lead)) if lead else 1
    jitter = None
    if rng is not None:
      jitter = torch.rand(self.num_levels, n, generator=rng, device=self.engine.device)
    res, hist = self.engine.forward(flat, rd, float(train_frac), jitter, bool(compute_extras), bool(zero_glo))
    L = self.num_levels
    for l in range(L):
      for d in (res[l], hist[l]):
        for k in list(d.keys()):
          d[k] = d[k].reshape(lead + tuple(d[k].shape[1:]))
      if compute_extras:
        nvis = self.config.vis_num_rays
        S = hist[l]['weights'].shape[-1]
        res[l]['ray_sdist'] = hist[l]['sdist'].reshape(-1, S + 1)[:nvis]
        res[l]['ray_weights'] = hist[l]['weights'].reshape(-1, S)[:nvis]
    if compute_extras:
      rgbs = hist[-1]['rgb'].reshape((-1,) + tuple(hist[-1]['rgb'].shape[-2:]))[:self.config.vis_num_rays]
      res[-1]['ray_rgbs'] = rgbs
      final_rgb = torch.sum(rgbs * res[-1]['ray_weights'][..., None], dim=-2)
      for l in range(L - 1):   # proposal levels show the final average colour (models.py:314-325)
        res[l]['ray_rgbs'] = final_rgb[:, None, :].expand(res[l]['ray_weights'].shape + (3,))
    return res, hist


def construct_model(rng, rays, config, max_rays: Optional[int] = None, device=None):
  """models.construct_model (models.py:333-357): returns (model, init_variables)."""
  del rays   # the reference traces 10 dummy rays to build shapes; shapes here come from the config
  model = Model(config, max_rays=max_rays, device=device)
  return model, model.init(rng)


def render_image(render_fn: Callable, rays: utils.Rays, rng, config, verbose: bool = True,
                 world_size: int = 1) -> Dict[str, Any]:
  """models.render_image (models.py:568-649): chunked full-frame render.

  render_fn(rng, chunk_rays) -> (renderings, ray_history); chunk_rays is sharded [world, n/world, C] like
  the reference's pmap input and the renderings carry a leading gathered axis (v[0] is taken).
  """
  height, width = rays.origins.shape[:2]
  num_rays = height * width
  flat = rays.map(lambda r: r.reshape(num_rays, -1))
  chunks = []
  idx0s = range(0, num_rays, config.render_chunk_size)
  for i_chunk, idx0 in enumerate(idx0s):
    if verbose and i_chunk % max(1, len(idx0s) // 10) == 0:
      print(f'Rendering chunk {i_chunk}/{len(idx0s) - 1}')
    chunk = flat.map(lambda r: r[idx0:idx0 + config.render_chunk_size])
    actual = chunk.origins.shape[0]
    rem = actual % world_size
    padding = world_size - rem if rem else 0
    if padding:
      chunk = chunk.map(lambda r: torch.cat([r, r[-1:].expand((padding,) + tuple(r.shape[1:]))], 0))
    chunk = chunk.map(lambda r: utils.shard(r, world_size))
    renderings, _ = render_fn(rng, chunk)
    renderings = [{k: (utils.unshard(v[0], padding) if not isinstance(v, list) else v) for k, v in r.items()}
                  for r in renderings]
    out = renderings[-1]
    for k in list(renderings[0].keys()):
      if k.startswith('ray_'):
        out[k] = [r[k] for r in renderings]
    chunks.append(out)
  rendering: Dict[str, Any] = {}
  for k in chunks[0]:
    if k.startswith('ray_'):
      rendering[k] = [torch.cat([c[k][l] for c in chunks]) for l in range(len(chunks[0][k]))]
    else:
      z = torch.cat([c[k] for c in chunks])
      rendering[k] = z.reshape((height, width) + tuple(z.shape[1:]))
  keys = [k for k in rendering if k.startswith('ray_')]
  if keys:
    n = rendering[keys[0]][0].shape[0]
    ray_idx = torch.randperm(n, generator=torch.Generator().manual_seed(0))[:config.vis_num_rays]
    for k in keys:
      rendering[k] = [r[ray_idx.to(r.device)] for r in rendering[k]]
  return rendering


def frame_stripe(height: int, rank: int, world: int):
  """Pixel rows [row0, row1) of rank `rank`: equal stripes of ceil(height / world) rows (the last ranks may get fewer or none),
  so that the all-gathered, zero-padded stripes concatenate to the frame.  Returns (rows per stripe, row0, row1)."""
  rows = -(-height // world)
  return rows, min(rank * rows, height), min((rank + 1) * rows, height)


def render_frame(model: 'Model', variables, dataset, cam_idx: int, train_frac: float, config,
                 compute_extras: bool = True, want_u8: bool = False, want_psnr: bool = False) -> Dict[str, Any]:
  """The full-frame pipeline of eval.py:104-160 / render.py:164-187 for one camera of a device-resident dataset
  (`datasets.DeviceDataset` or a `datasets.Dataset` over it): `generate_ray_batch` + `models.render_image`
  (models.py:568-649) + the metric / quantisation steps, without the host loop.

  Every rank renders its stripe of pixel rows with ONE library call (`hugs_render_frame`: on-device ray generation,
  chunking, frame-sized outputs) and the stripes are all-gathered once per frame (the reference gathers per chunk,
  train_utils.py:559).  Returns rgb [H, W, 3], acc [H, W], distance_mean / distance_median [H, W] (compute_extras) as
  device tensors like `render_image`, plus `rgb_u8` (utils.save_img_u8's bytes) and `psnr` / `psnr_quantized`
  (image.mse_to_psnr of the mean squared error against the dataset image; eval_quantize_metrics) on request.
  """
  import torch.distributed as dist
  dd = getattr(dataset, 'device_dataset', dataset)
  if not 0 <= int(cam_idx) < dd.n_cams:
    raise IndexError(f'render_frame: camera {cam_idx} outside [0, {dd.n_cams})')
  if want_psnr and dd.images is None and dd.images_u8 is None:
    raise ValueError('render_frame: want_psnr needs a dataset with images')
  h, w = int(dd.heights_np[cam_idx]), int(dd.widths_np[cam_idx])
  world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
  rank = dist.get_rank() if world > 1 else 0
  rows, row0, row1 = frame_stripe(h, rank, world)
  flat = variables if torch.is_tensor(variables) else model.flat_params(variables)
  model._ensure_packed(flat)
  out = model.engine.render_frame(flat, dd._cs, cam_idx, w, h, row0, row1, float(train_frac),
                                  zero_glo=config.enable_render_zero_glo, compute_extras=compute_extras, want_u8=want_u8,
                                  want_sse=want_psnr)
  sse = out.pop('sse', None)
  if world > 1:
    for k in list(out.keys()):
      v = out[k]
      pad = torch.zeros((rows,) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype)
      pad[:v.shape[0]] = v
      parts = [torch.empty_like(pad) for _ in range(world)]
      dist.all_gather(parts, pad)
      out[k] = torch.cat(parts)[:h]
    if sse is not None:
      dist.all_reduce(sse)
  if sse is not None:
    mse = (sse / float(3 * h * w)).cpu()
    out['psnr'] = float(-10.0 / math.log(10.0) * math.log(max(float(mse[0]), 1e-300)))
    out['psnr_quantized'] = float(-10.0 / math.log(10.0) * math.log(max(float(mse[1]), 1e-300)))
  return out
