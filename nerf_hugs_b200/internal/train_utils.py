"""train_utils.setup_model and the train / render step functions with the reference signatures
(MipNeRF360/internal/train_utils.py:372-596), one process per GPU.

The reference pmaps one program over the local devices and pmean's the gradient
(train_utils.py:457-459, 479-483).  Here every rank runs this module on its own GPU, the flat fp32
gradient is all-reduced with torch.distributed (NCCL over NVLink) and every rank applies the same
clip + Adam update (`hugs_adam_step`), so parameters stay replicated exactly like under pmap.
"""
import dataclasses
import math as _pymath
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist

from .. import _lib
from . import math as hmath
from . import models
from . import utils


@dataclasses.dataclass
class TrainState:
  """flax TrainState analogue: optimiser step, flat fp32 parameters and Adam moments (device tensors)."""
  step: int
  params: torch.Tensor
  mu: torch.Tensor
  nu: torch.Tensor

  def tree(self, model):
    """Parameters as the flax-named pytree {'params': {'NerfMLP_0': {'Dense_0': {'kernel','bias'}}, ...}}."""
    return {'params': model.engine.unflatten_params(self.params)}


def _world():
  return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def allreduce_sum_(tensors):
  """In-place all-reduce(SUM) of the flat gradient and the stats vector across ranks; together with
  grad_scale = 1/world in hugs_adam_step (and the 1/world of _LazyStats) this is jax.lax.pmean of
  train_utils.py:457-459.  Returns the world size.  No-op on a single process."""
  rank, world = _world()
  if world > 1:
    for t in tensors:
      dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return world


def loss_cfg_from(config, is_finetune: bool = False) -> '_lib.LossCfg':
  c = _lib.LossCfg()
  c.data_loss_type = {'charb': 0, 'mse': 1}[config.data_loss_type]
  c.charb_padding, c.data_loss_mult = config.charb_padding, config.data_loss_mult
  c.data_coarse_loss_mult = config.data_coarse_loss_mult
  c.interlevel_loss_mult = 0.0 if is_finetune else config.interlevel_loss_mult
  c.distortion_loss_mult = 0.0 if is_finetune else config.distortion_loss_mult
  c.use_static_mask = int(config.transient_type == 'withmask' and not is_finetune)   # train_utils.py:422-425
  c.withmask_transient_weight = float(config.withmask_transient_weight)
  c.disable_multiscale_loss = int(config.disable_multiscale_loss)
  return c


def create_optimizer(config, variables, model):
  """train_utils.create_optimizer (train_utils.py:487-512): Adam state + the lr schedule."""
  flat = model.flat_params(variables)
  lr_fn = lambda step: hmath.learning_rate_decay(step, config.lr_init, config.lr_final, config.max_steps,
                                                config.lr_delay_steps, config.lr_delay_mult)
  return TrainState(step=0, params=flat, mu=torch.zeros_like(flat), nu=torch.zeros_like(flat)), lr_fn


def _to_device(x, device):
  if x is None:
    return None
  if not torch.is_tensor(x):
    x = torch.as_tensor(x)
  return x.to(device, non_blocking=True)


def _rng_seed(rng):
  """Seed of the per-step sampling randomness: a torch.Generator, an int, or None."""
  if rng is None:
    return 0
  if isinstance(rng, torch.Generator):
    return int(rng.initial_seed()) or 1
  return int(rng) or 1


def create_train_step(model: models.Model, config, is_finetune: bool = False):
  """train_utils.create_train_step (train_utils.py:372-484).

  train_pstep(rng, state, batch, train_frac, inlier_thresholds) -> (state, stats, rng)
    rng:   a torch.Generator / int seed (or None when sampling should be deterministic); together with the step count it
           keys the counter-based jitter stream of the sampling kernel (hugs_set_train_rng): no RNG kernel per step
    batch: utils.Batch of this rank's rays (host or device tensors; host tensors are staged on a copy stream, so the
           transfer of batch i overlaps the GPU work of step i - 1)
  `state` is updated in place and returned (the reference donates it, train_utils.py:483).
  """
  eng = model.engine
  if getattr(config, 'weight_decay_mults', None):
    raise NotImplementedError(
        "Config.weight_decay_mults is set: the 'weight' loss term (train_utils.py:444-447) is not on this path "
        '(no shipped gin sets it); silently training without it would change the optimisation')
  lcfg = loss_cfg_from(config, is_finetune)
  lr_fn = lambda step: hmath.learning_rate_decay(step, config.lr_init, config.lr_final, config.max_steps,
                                                config.lr_delay_steps, config.lr_delay_mult)
  dev = eng.device
  # the flat gradient and the 16 stats share one buffer (train_utils.py:457-459 pmean's both); the 16-float offset keeps
  # the stats 64-byte aligned
  bucket = torch.zeros(eng.n_params + 16 + (-eng.n_params) % 16, device=dev)
  grad = bucket[:eng.n_params]
  stats_dev = bucket[bucket.numel() - 16:]
  norms_dev = torch.empty(9, device=dev)
  n_t = len(eng.layout)
  tstats_dev = torch.empty(n_t, 5, device=dev)
  names = [t[0] for t in eng.layout]
  L = model.num_levels
  # flat layout: [NerfMLP_0 | PropMLP_0 | GloEmbed_0]; the NerfMLP_0 part is final before the proposal levels' backward
  nerf_end = max([off + r * c for _, off, r, c, mod in eng.layout if mod == 0], default=0)
  # every step's stats are copied (asynchronously) into a pinned host ring, so a stats object stays valid after later
  # steps have been launched and reading it waits only for its own step
  ring = torch.empty(_STATS_RING, 25 + 5 * n_t).pin_memory()
  holders = [None] * _STATS_RING
  ov = {'side': None, 'event': None, 'copy': None, 'slots': [None] * _STAGE_SLOTS, 'i': 0}

  def stage(batch):
    """(rays dict, rgb, slot): host tensors are copied to rotating device buffers on a copy stream; a batch that already
    lives on the device passes through (slot None)."""
    items = dict(batch.rays.as_dict())
    items['rgb'] = batch.rgb
    items = {k: (v if (v is None or torch.is_tensor(v)) else torch.as_tensor(v)) for k, v in items.items()}
    if all(v is None or v.device == dev for v in items.values()):
      return {k: v for k, v in items.items() if k != 'rgb'}, items['rgb'], None
    if ov['copy'] is None:
      ov['copy'] = torch.cuda.Stream(dev)
    slot = ov['i'] % _STAGE_SLOTS
    ov['i'] += 1
    bufs, free_ev = ov['slots'][slot] or ({}, None)
    with torch.cuda.stream(ov['copy']):
      if free_ev is not None:
        ov['copy'].wait_event(free_ev)        # the step that last read this slot has finished
      for k, v in items.items():
        if v is None:
          bufs[k] = None
          continue
        buf = bufs.get(k)
        if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
          buf = torch.empty(v.shape, dtype=v.dtype, device=dev)
        buf.copy_(v, non_blocking=True)
        bufs[k] = buf
      done = torch.cuda.Event()
      done.record(ov['copy'])
    torch.cuda.current_stream(dev).wait_event(done)
    ov['slots'][slot] = [bufs, None]
    return {k: v for k, v in bufs.items() if k != 'rgb'}, bufs['rgb'], slot

  def train_step(rng, state: TrainState, batch: utils.Batch, train_frac, inlier_thresholds=None):
    del inlier_thresholds   # read by compute_robustnerf_loss only; transient_type='robustnerf' is refused by models.Model
    rank, world = _world()
    rays, rgb, slot_in = stage(batch)
    eng.set_train_rng(_rng_seed(rng) if config.randomized else 0, state.step)
    model._ensure_packed(state.params)
    if world > 1 and ov['event'] is None:
      ov['side'] = torch.cuda.Stream(dev)
      ov['event'] = torch.cuda.Event()
      ov['event'].record(torch.cuda.current_stream(dev))          # materialises the cudaEvent_t
      eng.set_grad_ready_event(ov['event'])
    eng.loss_and_grad(state.params, rays, rgb[..., :3], float(train_frac), None, lcfg, grad, stats_dev)
    if slot_in is not None:                                       # the staging slot is free once this step has read it
      ev = torch.cuda.Event()
      ev.record(torch.cuda.current_stream(dev))
      ov['slots'][slot_in][1] = ev
    if world > 1:
      # pmean(grad), pmean(stats) (train_utils.py:457-459) as two collectives on one communicator: the NerfMLP_0 part
      # (87 % of the floats) starts as soon as it is final, under the proposal levels' backward pass; the rest follows
      with torch.cuda.stream(ov['side']):
        ov['side'].wait_event(ov['event'])
        early = dist.all_reduce(bucket[:nerf_end], op=dist.ReduceOp.SUM, async_op=True)
      dist.all_reduce(bucket[nerf_end:], op=dist.ReduceOp.SUM)
      early.wait()
    a = _lib.AdamCfg()
    a.lr = float(lr_fn(state.step))
    a.beta1, a.beta2, a.eps = config.adam_beta1, config.adam_beta2, config.adam_eps
    a.grad_max_norm, a.grad_max_val = config.grad_max_norm, config.grad_max_val
    a.step, a.grad_scale = int(state.step), 1.0 / world
    eng.adam_step(state.params, grad, state.mu, state.nu, a, norms_dev, tstats_dev)
    model._packed_version = (state.params, state.params._version)   # adam_step re-packs the bf16 operands
    slot = state.step % _STATS_RING
    if holders[slot] is not None:
      holders[slot]._fetch()                          # about to reuse the slot: materialise its old owner
    ring[slot, :16].copy_(stats_dev, non_blocking=True)
    ring[slot, 16:25].copy_(norms_dev, non_blocking=True)
    ring[slot, 25:].copy_(tstats_dev.reshape(-1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    state.step += 1
    stats = _LazyStats(ring[slot], ev, L, world, a.lr, names)
    holders[slot] = stats
    return state, stats, rng

  return train_step


_STAGE_SLOTS = 3
_STATS_RING = 64


def _summarize(names, values, reduce):
  """train_utils.summarize_tree(tree, fn, max_depth=3): 'Module', 'Module/Dense_k' and 'Module/Dense_k/kernel' keys."""
  out = {}
  for name, v in zip(names, values):
    parts = name.split('/')
    for d in range(1, len(parts) + 1):
      k = '/'.join(parts[:d])
      out[k] = v if k not in out else reduce(out[k], v)
  return out


class _LazyStats(dict):
  """stats pytree of train_step (train_utils.py:442-476).  The values sit in a pinned host slot filled by an
  asynchronous copy; the first access (of any kind) waits for that step's event only, so the training loop does not
  synchronise every step (the reference reads them every print_every)."""

  def __init__(self, host_slot, event, L, world, lr, names):
    super().__init__()
    self._h, self._ev, self._L, self._world, self._lr, self._done = host_slot, event, L, world, lr, False
    self._names = names

  def _fetch(self):
    if self._done:
      return
    self._done = True
    if self._ev is not None:
      self._ev.synchronize()
    import numpy as np
    s = self._h[:16].numpy() / self._world
    t = self._h[25:].numpy().reshape(-1, 5).astype(np.float64)
    mses = np.array(s[4:4 + self._L])
    psnrs = -10.0 / _pymath.log(10.0) * np.log(np.maximum(mses, 1e-30))      # image.mse_to_psnr
    add, mx = (lambda a, b: a + b), max
    dict.update(self, {
        'loss': float(s[0]),
        'losses': {'data': float(s[1]), 'interlevel': float(s[2]), 'distortion': float(s[3])},
        'mses': mses, 'psnrs': psnrs, 'psnr': float(psnrs[-1]), 'lr': self._lr,
        'weight_l2s': _summarize(self._names, t[:, 0], add),                                   # tree_norm_sq
        'grad_norms': {k: _pymath.sqrt(v) for k, v in _summarize(self._names, t[:, 1], add).items()},
        'grad_maxes': _summarize(self._names, t[:, 2], mx),
        'opt_update_norms': {k: _pymath.sqrt(v) for k, v in _summarize(self._names, t[:, 3], add).items()},
        'opt_update_maxes': _summarize(self._names, t[:, 4], mx),
    })

  def __getitem__(self, k):
    self._fetch()
    return dict.__getitem__(self, k)

  def __contains__(self, k):
    self._fetch()
    return dict.__contains__(self, k)

  def __iter__(self):
    self._fetch()
    return dict.__iter__(self)

  def __len__(self):
    self._fetch()
    return dict.__len__(self)

  def get(self, k, default=None):
    self._fetch()
    return dict.get(self, k, default)

  def keys(self):
    self._fetch()
    return dict.keys(self)

  def values(self):
    self._fetch()
    return dict.values(self)

  def items(self):
    self._fetch()
    return dict.items(self)


def create_render_fn(model: models.Model, config):
  """train_utils.create_render_fn (train_utils.py:555-575).

  render_eval_pfn(variables, train_frac, _, rays) with rays sharded [world, n/world, C]; this rank renders
  its shard and the shards are all-gathered so the result carries the reference's leading gathered axis.
  """
  def render_eval_fn(variables, train_frac, _, rays: utils.Rays):
    rank, world = _world()
    mine = rays.map(lambda r: r[rank] if r.shape[0] == world else r[0])
    res, hist = model.apply(variables, None, mine, train_frac, True, zero_glo=config.enable_render_zero_glo)
    out = []
    for r in res:
      d = {}
      for k, v in r.items():
        if world > 1:
          parts = [torch.empty_like(v) for _ in range(world)]
          dist.all_gather(parts, v.contiguous())
          full = torch.stack(parts)
        else:
          full = v[None]
        d[k] = full[None]            # [1(gather axis taken as v[0]), world, n/world, ...]
      out.append(d)
    return out, hist

  return render_eval_fn


def setup_model(config, rng, max_rays: Optional[int] = None, device=None):
  """train_utils.setup_model (train_utils.py:579-596):
  returns (model, state, render_eval_pfn, train_pstep, lr_fn)."""
  model, variables = models.construct_model(rng, utils.dummy_rays(), config, max_rays=max_rays, device=device)
  state, lr_fn = create_optimizer(config, variables, model)
  render_eval_pfn = create_render_fn(model, config)
  train_pstep = create_train_step(model, config, False)
  return model, state, render_eval_pfn, train_pstep, lr_fn
