"""nerfacto (hash-grid field + proposal networks) of the torch twins (`/root/reference/nerfacto/models/nerfacto.py`) on the
B200 engine.

Same constructor, `forward(batch, curr_step, perturb, chunk_size)`, `get_params_dict()`, output keys (`rgb, depth,
accumulation, depth_prop_i, accumulation_prop_i`; training adds `weights_list, spacing_bins_list`, nerfacto.py:410-412)
and `state_dict()` names as the reference with `enable_tcnn_mlp: False` (every shipped yml).  Sampling, the hash-grid
encoding, all MLPs, compositing, the three losses and every backward pass run in libhugs_b200.so
(nerf_hugs_b200/nerfacto/ops.py); there is no torch fallback.

tiny-cuda-nn is not a dependency here: the encoder containers below own the `params` tensor in tcnn's layout and the
library restates the encoding (nerf_hugs_b200/csrc/hashfield.cu).

Not built (loud NotImplementedError): `enable_tcnn_mlp: True` (tcnn's fused-MLP parameter layout), the NeRF-W / HA-NeRF /
RobustNeRF heads and losses, MLP shapes other than the shipped 256 / 64 / 256 (field) and 64 (proposals).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import os

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import Parameter

from .. import ops
from ..utils.utils import merge_tensor_data, split_tensor_data
from ... import _lib


@dataclass
class ModelConfig:
  """Field set of the reference's dataclass (nerfacto.py:18-114)."""
  num_levels: int = 16
  base_res: int = 16
  max_res: int = 2048
  log2_hashmap_size: int = 19
  features_per_level: int = 2
  hidden_dim: int = 64
  geo_feat_dim: int = 15
  hidden_dim_color: int = 64
  hidden_dim_transient: int = 128
  density_activation: str = 'trunc_exp'
  enable_tcnn_mlp: bool = True
  beta_min: float = 0.1

  transient_type: Optional[str] = None
  num_embedding: int = 3500
  use_appearance_embedding: bool = False
  use_transient_embedding: bool = False
  appearance_embedding_dim: int = 32
  transient_embedding_dim: int = 16
  eval_embedding: str = 'average'

  num_levels_implicit: int = 8
  base_res_implicit: int = 16
  max_res_implicit: int = 1024
  log2_hashmap_size_implicit: int = 17
  features_per_level_implicit: int = 2
  hidden_dim_implicit: int = 128

  num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
  num_nerf_samples_per_ray: int = 48
  proposal_update_every: int = 5
  proposal_warmup: int = 5000
  num_proposal_iterations: int = 2
  use_same_proposal_network: bool = False
  proposal_net_args_list: List[Dict] = field(
      default_factory=lambda: [
          {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128},
          {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256},
      ]
  )
  proposal_initial_sampler: Optional[str] = None
  proposal_histogram_padding: float = 0.01
  use_proposal_weight_anneal: bool = True
  proposal_weights_anneal_slope: float = 10.0
  proposal_weights_anneal_max_num_iters: int = 1000
  use_single_jitter: bool = True
  opaque_background: bool = False

  rgb_loss_type: str = 'mse'
  rgb_charb_loss_padding: float = 0.001
  rgb_loss_mult: float = 1.0
  interlevel_loss_mult: float = 1.0
  distortion_loss_mult: float = 0.002
  nerfw_beta_loss_mult: float = 1.0
  nerfw_beta_loss_bias: float = 3.0
  nerfw_density_loss_mult: float = 0.01
  hanerf_mask_size_loss_mult_min: float = 6e-3
  hanerf_mask_size_loss_mult_max: float = 5e-2
  hanerf_mask_size_loss_mult_k: float = 1e-3
  robustnerf_inlier_quantile: float = 0.8
  robustnerf_smoothed_filter_size: int = 3
  robustnerf_smoothed_inlier_quantile: float = 0.5
  robustnerf_inner_patch_size: int = 8
  robustnerf_inner_patch_inlier_quantile: float = 0.4
  withmask_transient_weight: float = 0.


def _level_entries(n_levels, base_res, per_level_scale, log2_hashmap_size):
  """Total entries of a tcnn HashGrid (float32 arithmetic as in its GridEncoding constructor)."""
  l2s = np.log2(np.float32(per_level_scale)).astype(np.float32)
  total = 0
  for l in range(n_levels):
    scale = np.float32(np.exp2(np.float32(l) * l2s).astype(np.float32) * np.float32(base_res) - np.float32(1.0))
    res = int(np.ceil(scale)) + 1
    cnt = min(res ** 3, 0xFFFFFFFF // 2)
    cnt = min((cnt + 7) // 8 * 8, 1 << log2_hashmap_size)
    total += cnt
  return total


class GridEncoder(nn.Module):
  """Stands where `tcnn.Encoding` stands in the reference's module tree: owns `params` (tcnn's flat layout, uniform in
  [-1e-4, 1e-4] like tcnn's default initialisation) and the encoding's configuration.  otype 'SphericalHarmonics' has no
  parameters (an empty `params`, as tcnn's torch module registers)."""

  def __init__(self, otype: str, n_levels=0, base_res=0, per_level_scale=1.0, log2_hashmap_size=0, features_per_level=2):
    super().__init__()
    self.otype = otype
    if otype == 'HashGrid':
      self.n_levels, self.base_res, self.per_level_scale = n_levels, base_res, float(per_level_scale)
      self.log2_hashmap_size, self.features_per_level = log2_hashmap_size, features_per_level
      self.n_output_dims = n_levels * features_per_level
      total = _level_entries(n_levels, base_res, per_level_scale, log2_hashmap_size)
      self.params = nn.Parameter((torch.rand(total * features_per_level, dtype=torch.float32) * 2 - 1) * 1e-4)
    else:
      self.n_output_dims = 16
      self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32))

  def forward(self, *a, **k):
    raise RuntimeError('GridEncoder is a parameter container; Model.forward runs the encoding on the engine')


def _check_shapes(hidden_dim, ok, what):
  if hidden_dim != ok:
    raise NotImplementedError(f'{what} must be {ok} on the engine (every shipped yml), got {hidden_dim}')


class _HashFieldBase(nn.Module):
  def _register_grid_buffers(self, base_res, max_res, num_levels, log2_hashmap_size):
    self.register_buffer("base_res", torch.tensor(base_res))
    self.register_buffer("max_res", torch.tensor(max_res))
    self.register_buffer("num_levels", torch.tensor(num_levels))
    self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))

  def forward(self, *a, **k):
    raise RuntimeError('parameter container; Model.forward runs the field on the engine')


class NerfactoField(_HashFieldBase):
  """Module tree of nerfacto.py:643-809 with enable_tcnn_mlp=False: `direction_encoder`, `mlp_base` = [grid, Linear, ReLU,
  Linear], `mlp_head` = [Linear, ReLU, Linear, ReLU, Linear]."""

  def __init__(self, bound, num_levels, base_res, max_res, log2_hashmap_size, features_per_level, hidden_dim, geo_feat_dim,
               hidden_dim_color, density_activation, appearance_embedding_dim, contract):
    super().__init__()
    _check_shapes(hidden_dim, 256, 'hidden_dim'); _check_shapes(geo_feat_dim, 64, 'geo_feat_dim')
    _check_shapes(hidden_dim_color, 256, 'hidden_dim_color')
    self.bound, self.contract, self.geo_feat_dim = bound, contract, geo_feat_dim
    self.appearance_embedding_dim = appearance_embedding_dim
    self.density_bias, self.rgb_bias = -1., 0.
    self._register_grid_buffers(base_res, max_res, num_levels, log2_hashmap_size)
    self.direction_encoder = GridEncoder('SphericalHarmonics')
    growth = np.exp((np.log(max_res) - np.log(base_res)) / (num_levels - 1))
    grid = GridEncoder('HashGrid', num_levels, base_res, growth, log2_hashmap_size, features_per_level)
    l1 = nn.Linear(grid.n_output_dims, hidden_dim); torch.nn.init.kaiming_uniform_(l1.weight)
    l2 = nn.Linear(hidden_dim, 1 + geo_feat_dim); torch.nn.init.kaiming_uniform_(l2.weight)
    self.mlp_base = nn.Sequential(grid, l1, nn.ReLU(), l2)
    in_dim = 16 + geo_feat_dim + appearance_embedding_dim
    h1 = nn.Linear(in_dim, hidden_dim_color); torch.nn.init.kaiming_uniform_(h1.weight)
    h2 = nn.Linear(hidden_dim_color, hidden_dim_color); torch.nn.init.kaiming_uniform_(h2.weight)
    h3 = nn.Linear(hidden_dim_color, 3); torch.nn.init.kaiming_uniform_(h3.weight)
    self.mlp_head = nn.Sequential(h1, nn.ReLU(), h2, nn.ReLU(), h3)

  def grid(self) -> GridEncoder:
    return self.mlp_base[0]

  def entries(self, embedding: Optional[nn.Embedding]):
    """(flat tensor[#row0:rows], torch tensor, element offset, transpose, ld) - see hugs_tensor_copy / hugs_hashfield_layout"""
    l1, l2 = self.mlp_base[1], self.mlp_base[3]
    h1, h2, h3 = self.mlp_head[0], self.mlp_head[2], self.mlp_head[4]
    G, app = self.geo_feat_dim, self.appearance_embedding_dim
    H = l2.in_features
    e = [('base0/kernel', l1.weight, 0, 1, 0), ('base0/bias', l1.bias, 0, 0, 0),
         ('density/kernel', l2.weight, 0, 1, H), ('density/bias', l2.bias, 0, 0, 0),
         ('geo/kernel', l2.weight, H, 1, H), ('geo/bias', l2.bias, 1, 0, 0),
         # the colour MLP reads [SH | geometry | appearance] (nerfacto.py:851-854); the engine keeps geometry rows first
         (f'head0/kernel#0:{G}', h1.weight, 16, 1, h1.in_features), (f'head0/kernel#{G}:16', h1.weight, 0, 1, h1.in_features)]
    if app > 0:
      e.append((f'head0/kernel#{G + 16}:{app}', h1.weight, 16 + G, 1, h1.in_features))
    e += [('head0/bias', h1.bias, 0, 0, 0), ('head1/kernel', h2.weight, 0, 1, 0), ('head1/bias', h2.bias, 0, 0, 0),
          ('rgb/kernel', h3.weight, 0, 1, 0), ('rgb/bias', h3.bias, 0, 0, 0)]
    if app > 0:
      e.append(('embedding', embedding.weight, 0, 0, 0))
    return e


class HashMLPDensityField(_HashFieldBase):
  """Module tree of nerfacto.py:878-989 with enable_tcnn_mlp=False: `mlp_base` = [grid, Linear, ReLU, Linear]."""

  def __init__(self, bound, density_activation, contract, num_levels=8, base_res=16, max_res=1024, log2_hashmap_size=18,
               features_per_level=2, num_layers=2, hidden_dim=64):
    super().__init__()
    _check_shapes(hidden_dim, 64, 'proposal hidden_dim'); _check_shapes(num_layers, 2, 'proposal num_layers')
    self.bound, self.contract = bound, contract
    self.density_bias = -1.
    self._register_grid_buffers(base_res, max_res, num_levels, log2_hashmap_size)
    growth = np.exp((np.log(max_res) - np.log(base_res)) / (num_levels - 1))
    grid = GridEncoder('HashGrid', num_levels, base_res, growth, log2_hashmap_size, features_per_level)
    l1 = nn.Linear(grid.n_output_dims, hidden_dim); torch.nn.init.kaiming_uniform_(l1.weight)
    l2 = nn.Linear(hidden_dim, 1); torch.nn.init.kaiming_uniform_(l2.weight)
    self.mlp_base = nn.Sequential(grid, l1, nn.ReLU(), l2)

  def grid(self) -> GridEncoder:
    return self.mlp_base[0]

  def entries(self, embedding=None):
    l1, l2 = self.mlp_base[1], self.mlp_base[3]
    return [('base0/kernel', l1.weight, 0, 1, 0), ('base0/bias', l1.bias, 0, 0, 0),
            ('density/kernel', l2.weight, 0, 1, 0), ('density/bias', l2.bias, 0, 0, 0)]


def _distinct(entries):
  out, seen = [], set()
  for _, t, _, _, _ in entries:
    if id(t) not in seen:
      seen.add(id(t)); out.append(t)
  return out


class Model(nn.Module):
  def __init__(self, config: ModelConfig, bound: Optional[float], enable_amp: bool, enable_scene_contraction: bool) -> None:
    super().__init__()
    self.config = config
    self.bound = bound
    self.enable_amp = enable_amp              # the engine's MLPs always run bf16 x bf16 -> fp32 on the tensor cores
    self.enable_scene_contraction = enable_scene_contraction
    c = config
    assert self.bound is not None
    if enable_scene_contraction:
      assert self.bound == 2.0, f"When using scene contraction, bound should be set to 2, but got {self.bound}"
    if c.transient_type in ('nerfw', 'hanerf', 'robustnerf') or c.use_transient_embedding:
      raise NotImplementedError(f"transient_type={c.transient_type!r}: NeRF-W / HA-NeRF / RobustNeRF are out of scope (SURVEY.md §2.1)")
    if c.enable_tcnn_mlp:
      raise NotImplementedError('enable_tcnn_mlp: True (tcnn fused-MLP parameter layout) is not built; every shipped yml sets False')
    if c.density_activation not in ('trunc_exp', 'softplus'):
      raise NotImplementedError()
    if c.features_per_level != 2:
      raise NotImplementedError('features_per_level must be 2')
    if c.proposal_initial_sampler not in ops.SPACING:
      raise ValueError(f"Sampler does not support {c.proposal_initial_sampler}. ")

    self.use_appearance_embedding = c.use_appearance_embedding
    if self.use_appearance_embedding:
      self.embedding_appearance = nn.Embedding(c.num_embedding, c.appearance_embedding_dim)
      app_dim = c.appearance_embedding_dim
    else:
      self.embedding_appearance = None
      app_dim = 0
    self.use_transient_embedding = False
    self.embedding_transient = None
    self.implicit_mask = None

    self.field = NerfactoField(
        bound=self.bound, num_levels=c.num_levels, base_res=c.base_res, max_res=c.max_res,
        log2_hashmap_size=c.log2_hashmap_size, features_per_level=c.features_per_level, hidden_dim=c.hidden_dim,
        geo_feat_dim=c.geo_feat_dim, hidden_dim_color=c.hidden_dim_color, density_activation=c.density_activation,
        appearance_embedding_dim=app_dim, contract=enable_scene_contraction)
    self.proposal_networks = nn.ModuleList()
    if c.use_same_proposal_network:
      assert len(c.proposal_net_args_list) == 1, "Only one proposal network is allowed."
      nets = [c.proposal_net_args_list[0]]
    else:
      nets = [c.proposal_net_args_list[min(i, len(c.proposal_net_args_list) - 1)] for i in range(c.num_proposal_iterations)]
    for args in nets:
      self.proposal_networks.append(HashMLPDensityField(bound=self.bound, density_activation=c.density_activation,
                                                        contract=enable_scene_contraction, **args))
    bias = -1. if c.density_activation == 'softplus' else 0.      # trunc_exp takes the raw density as is (nerfacto.py:702-710)
    self._render_cfg = ops.render_cfg(c.opaque_background, c.density_activation, bias, 1., 0., 0.)
    # 'bf16_tc' (throughput) | 'tc_split' (fp32-level parity: bf16 hi + lo operands through the same tcgen05 GEMMs)
    self.precision = os.environ.get('HUGS_NERFACTO_PRECISION', 'bf16_tc')
    if self.precision not in ('bf16_tc', 'tc_split'):
      raise NotImplementedError(f"nerfacto twin: precision {self.precision!r} (use 'bf16_tc' or 'tc_split')")
    self._engines: Dict[str, ops.HashFieldEngine] = {}
    self.jitter_override = None     # test hook: list of draws per level, used in place of torch.rand (ray_utils.py:151-152)

  def get_params_dict(self) -> Dict[str, List[Parameter]]:
    params_dict = {'field': list(self.field.parameters()), 'proposal': list(self.proposal_networks.parameters())}
    if self.embedding_appearance is not None:
      params_dict['appearance_embedding'] = list(self.embedding_appearance.parameters())
    return params_dict

  # ---- engines -----------------------------------------------------------------------------------------------------
  def _engine(self, key: str, module, n_rays: int, n_samples: int, device) -> ops.HashFieldEngine:
    fe = self._engines.get(key)
    need = n_rays * n_samples
    # 'bf16_tc' (throughput) | 'tc_split' (fp32-level parity through the same tcgen05 GEMMs, forward and backward)
    precision = {'bf16_tc': 1, 'tc_split': 2}[self.precision]
    if (fe is not None and fe.device == torch.device(device) and fe.desc.max_samples >= need and fe.desc.max_rays >= n_rays
        and fe.desc.precision == precision):
      return fe
    g = module.grid()
    d = _lib.HashFieldDesc()
    d.n_levels, d.features_per_level, d.log2_hashmap_size, d.base_res = g.n_levels, g.features_per_level, g.log2_hashmap_size, g.base_res
    d.per_level_scale = g.per_level_scale
    if isinstance(module, NerfactoField):
      d.hidden_dim, d.geo_feat_dim, d.hidden_dim_color = 256, 64, 256
      d.appearance_dim = module.appearance_embedding_dim
      d.num_embeddings = self.config.num_embedding if module.appearance_embedding_dim > 0 else 0
    else:
      d.hidden_dim, d.geo_feat_dim, d.hidden_dim_color = 64, 0, 0
    d.bound, d.contract = float(self.bound), int(self.enable_scene_contraction)
    d.max_samples = max(need, fe.desc.max_samples if fe is not None else 0)
    d.max_rays = max(n_rays, fe.desc.max_rays if fe is not None else 0)
    d.precision = precision
    if fe is not None:
      fe.close()
    fe = ops.HashFieldEngine(d, device, g.params, module.entries(self.embedding_appearance))
    self._engines[key] = fe
    return fe

  def _embedding_override(self):
    """({id(embedding weight): stand-in}, zero_app) following get_embedding (nerfacto.py:266-284)."""
    if self.embedding_appearance is None:
      return None, False
    w = self.embedding_appearance.weight
    if self.training or self.config.eval_embedding == 'original':
      return None, False
    if self.config.eval_embedding == 'average':
      return {id(w): w.detach().mean(dim=0, keepdim=True).expand_as(w).contiguous()}, False
    if self.config.eval_embedding == 'zero':
      return None, True
    raise NotImplementedError(f"{self.config.eval_embedding} is not supported.")

  # ---- the reference's forward -------------------------------------------------------------------------------------
  def forward_rays(self, rays: Dict[str, Tensor], curr_step: int, perturb: bool) -> dict:
    c = self.config
    dev = rays['origin'].device
    if dev.type != 'cuda':
      raise RuntimeError('nerf_hugs_b200 runs on a CUDA device only: move the model and the batch to the GPU')
    n = rays['origin'].shape[0]
    near, far = rays['near'], rays['far']
    if c.use_proposal_weight_anneal:
      N = c.proposal_weights_anneal_max_num_iters
      train_frac = np.clip(curr_step / N, 0, 1)
      s = c.proposal_weights_anneal_slope
      anneal = (s * train_frac) / ((s - 1) * train_frac + 1)
    else:
      anneal = 1.0
    proposal_update_interval = int(np.clip(np.interp(curr_step, [0, c.proposal_warmup], [0, c.proposal_update_every]),
                                           1, c.proposal_update_every))
    enable_proposal_update = ((curr_step % proposal_update_interval) == 0)

    weights_list, spacing_bins_list = [], []
    spacing_bins = torch.cat([torch.zeros_like(near), torch.ones_like(far)], dim=-1).float()
    weights = torch.ones_like(near).float()
    domain = (0., 1.)
    outputs = {}
    grad_on = torch.is_grad_enabled()
    override, zero_app = self._embedding_override()
    for i_level in range(c.num_proposal_iterations + 1):
      is_prop = i_level < c.num_proposal_iterations
      ns = c.num_proposal_samples_per_ray[i_level] if is_prop else c.num_nerf_samples_per_ray
      jit = None if self.jitter_override is None else self.jitter_override[i_level]
      spacing_bins, euclidean_bins = ops.sample_intervals(spacing_bins, weights, anneal, c.proposal_histogram_padding, ns,
                                                          perturb, c.use_single_jitter, domain, c.proposal_initial_sampler,
                                                          near, far, jitter=jit)
      eng_rays = {'origins': rays['origin'], 'directions': rays['direction']}
      if is_prop:
        net_idx = 0 if c.use_same_proposal_network else i_level
        module = self.proposal_networks[net_idx]
        fe = self._engine(f'prop{net_idx}:{i_level}', module, n, ns, dev)
        with torch.set_grad_enabled(grad_on and enable_proposal_update):
          training = self.training and torch.is_grad_enabled()
          depth, acc, w = ops.render_hash_field(fe, eng_rays, euclidean_bins, None, self._render_cfg, training, False,
                                                module.grid().params, _distinct(fe.entries))
        color = None
      else:
        eng_rays['viewdirs'] = rays['viewdir']
        if self.use_appearance_embedding:
          eng_rays['embed_idx'] = rays['embed_idx']
        fe = self._engine('field', self.field, n, ns, dev)
        training = self.training and torch.is_grad_enabled()
        color, depth, acc, w = ops.render_hash_field(fe, eng_rays, euclidean_bins, rays['bg_rgb'], self._render_cfg, training,
                                                     zero_app, self.field.grid().params, _distinct(fe.entries), override)
      weights = w
      weights_list.append(w)
      spacing_bins_list.append(spacing_bins)
      suffix = f'_prop_{i_level}' if is_prop else ''
      if color is not None:
        outputs[f'rgb{suffix}'] = color
      outputs[f'depth{suffix}'] = depth
      outputs[f'accumulation{suffix}'] = acc
    if self.training:
      outputs['weights_list'] = weights_list
      outputs['spacing_bins_list'] = spacing_bins_list
    return outputs

  def forward(self, batch: Dict[str, Tensor], curr_step: int, perturb: bool, chunk_size: Optional[int] = None) -> dict:
    if self.training:
      outputs = self.forward_rays(batch, curr_step, perturb)
    else:
      batch_list = split_tensor_data(batch, chunk_size)
      outputs_list = [self.forward_rays(sub_batch, curr_step, perturb) for sub_batch in batch_list]
      outputs = merge_tensor_data(outputs_list)
    return outputs


class Loss(nn.Module):
  """criterion_dict['nerfacto'] (nerfacto.py:428-640): photometric loss (optionally HuGS-mask weighted) + interlevel +
  distortion."""

  def __init__(self, model: Model) -> None:
    super().__init__()
    self.config = model.config
    if self.config.rgb_loss_type not in ops.LOSS_TYPE:
      raise NotImplementedError()

  def _data_loss(self, outputs, batch, extra_infos, static_mask):
    c = self.config
    rgb_loss, mse = ops.rgb_loss(outputs['rgb'], batch['rgb'], static_mask, c.withmask_transient_weight, c.rgb_loss_type,
                                 c.rgb_charb_loss_padding, c.rgb_loss_mult)
    return rgb_loss, {'rgb_loss': rgb_loss.detach(), 'mse': mse}, extra_infos

  def compute_data_loss(self, outputs, batch, data_shape, extra_infos):
    return self._data_loss(outputs, batch, extra_infos, None)

  def compute_withmask_loss(self, outputs, batch, data_shape, extra_infos):
    # the HuGS static-mask gather into the photometric loss (nerfacto.py:467-490)
    return self._data_loss(outputs, batch, extra_infos, batch['static_mask'])

  def forward(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor], data_shape, is_finetune: bool, extra_infos: dict):
    c = self.config
    if is_finetune or c.transient_type is None:
      loss, info_dict, extra_infos = self.compute_data_loss(outputs, batch, data_shape, extra_infos)
    elif c.transient_type == 'withmask':
      loss, info_dict, extra_infos = self.compute_withmask_loss(outputs, batch, data_shape, extra_infos)
    else:
      raise NotImplementedError()
    if c.interlevel_loss_mult > 0:
      interlevel_loss_ = c.interlevel_loss_mult * ops.interlevel_loss(outputs['weights_list'], outputs['spacing_bins_list'])
      loss = loss + interlevel_loss_
      info_dict['interlevel_loss'] = interlevel_loss_.detach()
    if c.distortion_loss_mult > 0:
      distortion_loss_ = c.distortion_loss_mult * ops.distortion_loss(outputs['weights_list'], outputs['spacing_bins_list'])
      loss = loss + distortion_loss_
      info_dict['distortion_loss'] = distortion_loss_.detach()
    return loss, info_dict, extra_infos
