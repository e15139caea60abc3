"""Vanilla NeRF of the torch twins (`/root/reference/nerfacto/models/nerf.py`) on the B200 engine.

Same constructor, `forward(batch, curr_step, perturb, chunk_size)`, `get_params_dict()`, output keys and `state_dict()`
names as the reference (`nerf.py:119-126,228-241,263-383`), so that `nerfacto/train.py` / `eval.py` run unchanged; the
numerics - sampling, point encoding, the MLPs (tcgen05 chain kernel), compositing, the photometric loss and every
backward pass - run in libhugs_b200.so (nerf_hugs_b200/nerfacto/ops.py).  There is no torch fallback: a forward on CPU
tensors raises.

Not built (loud NotImplementedError): the NeRF-W / HA-NeRF / RobustNeRF heads and losses (out of scope),
`net_activation != 'relu'`, widths other than 256 / 256 / 128, view / transient sub-MLPs deeper than one layer, noise.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import Parameter

from .. import ops
from ..utils.utils import merge_tensor_data, split_tensor_data
from ...engine import EngineConfig


@dataclass
class ModelConfig:
  """Field set of the reference's dataclass (nerf.py:16-116); unsupported values fail in Model.__init__."""
  net_depth: int = 8
  net_width: int = 256
  bottleneck_width: int = 256
  net_depth_viewdirs: int = 1
  net_width_viewdirs: int = 128
  net_depth_transient: int = 4
  net_width_transient: int = 128
  net_activation: str = 'relu'
  min_deg_point: int = 0
  max_deg_point: int = 12
  skip_layer: int = 4
  skip_layer_dir: int = 4
  skip_layer_transient: int = 4
  deg_view: int = 4
  bottleneck_noise: float = 0.0
  density_activation: str = 'softplus'
  density_bias: float = -1.
  density_noise: float = 0.
  rgb_premultiplier: float = 1.
  rgb_activation: str = 'sigmoid'
  rgb_bias: float = 0.
  rgb_padding: float = 0.001
  beta_min: float = 0.1

  transient_type: Optional[str] = None
  num_embedding: int = 3500
  use_appearance_embedding: bool = False
  use_transient_embedding: bool = False
  appearance_embedding_dim: int = 32
  transient_embedding_dim: int = 16
  eval_embedding: str = 'average'

  net_depth_implicit: int = 4
  net_width_implicit: int = 256
  deg_implicit: int = 10

  num_coarse_nerf_samples_per_ray: int = 64
  num_fine_nerf_samples_per_ray: int = 128
  proposal_initial_sampler: str = 'uniform'
  use_single_jitter: bool = False
  opaque_background: bool = False

  rgb_loss_type: str = 'mse'
  rgb_charb_loss_padding: float = 0.001
  coarse_rgb_loss_mult: float = 1.0
  fine_rgb_loss_mult: float = 1.0

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


class MLP(nn.Module):
  """Parameter container with the reference's module tree (nerf.py:632-786): `mlp_base.{i}.{2j}`, `mlp_density`,
  `mlp_bottleneck`, `mlp_head.0.0`, `mlp_rgb`, created in the same order from the same RNG draws, so that a seed gives the
  reference's initial weights and its checkpoints load by name.  The arithmetic lives in the engine, not here."""

  def __init__(self, net_depth, net_width, bottleneck_width, appearance_embedding_dim, net_width_viewdirs, skip_layer,
               min_deg_point, max_deg_point, deg_view):
    super().__init__()
    self.net_depth, self.net_width, self.skip_layer = net_depth, net_width, skip_layer
    torch.rand((1, 3), dtype=torch.float32)          # the reference probes pos_enc with a random point (nerf.py:705)
    in_dim = 3 + 6 * (max_deg_point - min_deg_point)
    mlp_base, sub_mlp, last_dim = [], [], in_dim
    for i in range(net_depth):
      lin = nn.Linear(last_dim, net_width)
      torch.nn.init.kaiming_uniform_(lin.weight)
      sub_mlp += [lin, nn.ReLU()]
      if i % skip_layer == 0 and i > 0:
        last_dim = net_width + in_dim
        mlp_base.append(nn.Sequential(*sub_mlp))
        sub_mlp = []
      else:
        last_dim = net_width
    if len(sub_mlp) > 0:
      mlp_base.append(nn.Sequential(*sub_mlp))
    self.mlp_base = nn.ModuleList(mlp_base)
    self.mlp_density = nn.Linear(last_dim, 1)
    torch.nn.init.kaiming_uniform_(self.mlp_density.weight)
    self.mlp_bottleneck = nn.Linear(net_width, bottleneck_width)
    torch.rand((1, 3), dtype=torch.float32)          # nerf.py:735
    in_dim = bottleneck_width + 3 + 6 * deg_view + appearance_embedding_dim
    lin = nn.Linear(in_dim, net_width_viewdirs)
    torch.nn.init.kaiming_uniform_(lin.weight)
    self.mlp_head = nn.ModuleList([nn.Sequential(lin, nn.ReLU())])
    self.mlp_rgb = nn.Linear(net_width_viewdirs, 3)
    torch.nn.init.kaiming_uniform_(self.mlp_rgb.weight)

  def linears(self) -> List[nn.Linear]:
    """Dense layers in the engine's (flax creation) order: trunk, density, bottleneck, view, rgb."""
    trunk = [m for seq in self.mlp_base for m in seq if isinstance(m, nn.Linear)]
    return trunk + [self.mlp_density, self.mlp_bottleneck, self.mlp_head[0][0], self.mlp_rgb]

  def forward(self, *args, **kwargs):
    raise RuntimeError('MLP is a parameter container; Model.forward runs the field on the engine')


_RAY_KEYS = {'origin': 'origins', 'direction': 'directions', 'viewdir': 'viewdirs'}


class Model(nn.Module):
  def __init__(self, config: ModelConfig, bound: Optional[float], enable_amp: bool, enable_scene_contraction: bool) -> None:
    super().__init__()
    self.config = config
    self.bound = bound                        # not used in nerf
    self.enable_amp = enable_amp              # the engine's bf16 tensor-core mode is always on; autocast has no effect on it
    self.enable_scene_contraction = enable_scene_contraction
    c = config
    if c.transient_type in ('nerfw', 'hanerf') or c.use_transient_embedding:
      raise NotImplementedError(f"transient_type={c.transient_type!r}: the NeRF-W / HA-NeRF heads are out of scope (SURVEY.md §2.1)")
    if c.net_activation != 'relu' or c.rgb_activation != 'sigmoid':
      raise NotImplementedError('only net_activation="relu" / rgb_activation="sigmoid" run on the tensor-core path')
    if (c.net_width, c.bottleneck_width, c.net_width_viewdirs, c.net_depth_viewdirs) != (256, 256, 128, 1):
      raise NotImplementedError('the chain kernel runs net_width 256, bottleneck 256, one 128-wide view layer '
                                f'(got {c.net_width}, {c.bottleneck_width}, {c.net_width_viewdirs} x {c.net_depth_viewdirs})')
    if c.bottleneck_noise != 0 or c.density_noise != 0:
      raise NotImplementedError('bottleneck_noise / density_noise are not supported (0 in every shipped config)')
    if c.density_activation not in ops.DENSITY_ACT:
      raise NotImplementedError(f'density_activation={c.density_activation!r}')
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
    arg = dict(net_depth=c.net_depth, net_width=c.net_width, bottleneck_width=c.bottleneck_width,
               appearance_embedding_dim=app_dim, net_width_viewdirs=c.net_width_viewdirs, skip_layer=c.skip_layer,
               min_deg_point=c.min_deg_point, max_deg_point=c.max_deg_point, deg_view=c.deg_view)
    self.field = nn.ModuleDict({'coarse': MLP(**arg), 'fine': MLP(**arg)})
    # 'bf16_tc' (throughput) | 'tc_split' (fp32-level parity through the same tensor-core kernels) | 'fp32' (render only)
    self.precision = os.environ.get('HUGS_NERFACTO_PRECISION', 'bf16_tc')
    self._engines: Dict[str, ops.FieldEngine] = {}
    self._render_cfg = ops.render_cfg(c.opaque_background, c.density_activation, c.density_bias, c.rgb_premultiplier,
                                      c.rgb_bias, c.rgb_padding)
    self.jitter_override = None     # test hook: {field_type: draws} used in place of torch.rand (ray_utils.py:151-152)
    self.last_bins = None           # test hook: set to {} to record the euclidean fenceposts of every field evaluation
    self.bins_override = None       # test hook: {field_type: euclidean fenceposts} evaluated instead of the sampled ones

  # ---- the reference's surface ------------------------------------------------------------------------------------
  def get_params_dict(self) -> Dict[str, List[Parameter]]:
    params_dict = {'field': list(self.field.parameters())}
    if self.embedding_appearance is not None:
      params_dict['appearance_embedding'] = list(self.embedding_appearance.parameters())
    return params_dict

  def num_samples(self, field_type: str) -> int:
    c = self.config
    if field_type == 'coarse':
      return c.num_coarse_nerf_samples_per_ray
    return c.num_coarse_nerf_samples_per_ray + c.num_fine_nerf_samples_per_ray

  def _field_engine(self, field_type: str, n_rays: int, device) -> ops.FieldEngine:
    fe = self._engines.get(field_type)
    if fe is not None and fe.device == torch.device(device) and fe.ecfg.max_rays >= n_rays and fe.ecfg.precision == self.precision:
      return fe
    c = self.config
    ecfg = EngineConfig(
        num_levels=1, num_nerf_samples=self.num_samples(field_type), nerf_depth=c.net_depth, nerf_width=c.net_width,
        bottleneck_width=c.bottleneck_width, view_width=c.net_width_viewdirs, skip_layer=c.skip_layer,
        min_deg_point=c.min_deg_point, max_deg_point=c.max_deg_point, deg_view=c.deg_view,
        nerf_contract=self.enable_scene_contraction, opaque_background=c.opaque_background,
        num_glo_features=c.appearance_embedding_dim if self.use_appearance_embedding else 0,
        num_embeddings=c.num_embedding, density_bias=c.density_bias, rgb_premultiplier=c.rgb_premultiplier,
        rgb_bias=c.rgb_bias, rgb_padding=c.rgb_padding, precision=self.precision,
        max_rays=max(n_rays, fe.ecfg.max_rays if fe is not None else 0), encoding='point_pe')
    if fe is not None:
      fe.engine.close()
    fe = ops.FieldEngine(ecfg, device, self.field[field_type].linears(), self.embedding_appearance)
    self._engines[field_type] = fe
    return fe

  def _embedding_for_call(self):
    """(tensor standing in for embedding_appearance.weight, zero_glo) following get_embedding (nerf.py:243-261)."""
    if self.embedding_appearance is None:
      return None, False
    w = self.embedding_appearance.weight
    if self.training or self.config.eval_embedding == 'original':
      return w, False
    if self.config.eval_embedding == 'average':
      return w.detach().mean(dim=0, keepdim=True).expand_as(w).contiguous(), False
    if self.config.eval_embedding == 'zero':
      return w, True
    raise NotImplementedError(f"{self.config.eval_embedding} is not supported.")

  def forward_rays(self, rays: Dict[str, Tensor], curr_step: int, perturb: bool) -> dict:
    c = self.config
    dev = rays['origin'].device
    if dev.type != 'cuda':
      raise RuntimeError('nerf_hugs_b200 runs on a CUDA device only: move the model and the batch to the GPU')
    n = rays['origin'].shape[0]
    near, far = rays['near'], rays['far']
    eng_rays = {v: rays[k] for k, v in _RAY_KEYS.items()}
    if self.use_appearance_embedding:
      eng_rays['embed_idx'] = rays['embed_idx']
    spacing_bins = torch.cat([torch.zeros_like(near), torch.ones_like(far)], dim=-1).float()
    weights = torch.ones_like(near).float()
    domain = (0., 1.)
    fn = c.proposal_initial_sampler
    emb, zero_glo = self._embedding_for_call()
    training = self.training and torch.is_grad_enabled()
    outputs = {}
    for field_type in ['coarse', 'fine']:
      ns = c.num_coarse_nerf_samples_per_ray if field_type == 'coarse' else c.num_fine_nerf_samples_per_ray
      jit = None if self.jitter_override is None else self.jitter_override.get(field_type)
      if field_type == 'coarse':
        spacing_bins, euclidean_bins = ops.sample_intervals(spacing_bins, weights, 1., 0., ns, perturb, c.use_single_jitter,
                                                            domain, fn, near, far, jitter=jit)
      else:
        bins_, _ = ops.sample_intervals(spacing_bins, weights, 1., 0., ns, perturb, c.use_single_jitter, domain, jitter=jit)
        spacing_bins, euclidean_bins = ops.merge_bins(spacing_bins, bins_, domain, fn, near, far)
      if self.bins_override is not None and field_type in self.bins_override:
        euclidean_bins = self.bins_override[field_type].to(dev).float().contiguous()
      if self.last_bins is not None:
        self.last_bins.setdefault(field_type, []).append(euclidean_bins)
      fe = self._field_engine(field_type, n, dev)
      params = list(fe.params)
      if emb is not None:
        params[-1] = emb
      color, depth, acc, w = ops.render_field(fe, eng_rays, euclidean_bins, rays.get('bg_rgb'), self._render_cfg,
                                              training, zero_glo, params)
      weights = w.detach()
      suffix = '' if field_type == 'fine' else f'_{field_type}'
      outputs[f'rgb{suffix}'] = color
      outputs[f'depth{suffix}'] = depth
      outputs[f'accumulation{suffix}'] = acc
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
  """criterion_dict['nerf'] (nerf.py:386-629): photometric loss of the coarse and the fine rendering."""

  def __init__(self, model: Model) -> None:
    super().__init__()
    self.config = model.config
    if self.config.rgb_loss_type not in ops.LOSS_TYPE:
      raise NotImplementedError()

  def _data_loss(self, outputs, batch, data_shape, extra_infos, static_mask):
    c = self.config
    loss_list, info_dict = [], {}
    gt_rgb = batch['rgb']
    for field_type in ['coarse', 'fine']:
      suffix = '' if field_type == 'fine' else f'_{field_type}'
      mult = c.coarse_rgb_loss_mult if field_type == 'coarse' else c.fine_rgb_loss_mult
      rgb_loss, mse = ops.rgb_loss(outputs[f'rgb{suffix}'], gt_rgb, static_mask, c.withmask_transient_weight,
                                   c.rgb_loss_type, c.rgb_charb_loss_padding, mult)
      loss_list.append(rgb_loss)
      info_dict[f'rgb_loss{suffix}'] = rgb_loss.detach()
      info_dict[f'mse{suffix}'] = mse
    return sum(loss_list), info_dict, extra_infos

  def compute_data_loss(self, outputs, batch, data_shape, extra_infos):
    return self._data_loss(outputs, batch, data_shape, extra_infos, None)

  def compute_withmask_loss(self, outputs, batch, data_shape, extra_infos):
    return self._data_loss(outputs, batch, data_shape, extra_infos, batch['static_mask'])

  def forward(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor], data_shape, is_finetune: bool, extra_infos: dict):
    if is_finetune or self.config.transient_type is None:
      return self.compute_data_loss(outputs, batch, data_shape, extra_infos)
    # quirk B4: the reference never dispatches compute_withmask_loss for the vanilla NeRF either (nerf.py:610-627)
    raise NotImplementedError()
