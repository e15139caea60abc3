"""Experiment configuration with the reference's field names (MipNeRF360/internal/configs.py:45-184)
and a minimal parser for the gin subset the shipped .gin files use.

gin itself is not a dependency.  The 19 shipped configs (MipNeRF360/configs/*.gin) and the
`--gin_bindings` strings of scripts/*.sh only use
    `Scope.field = literal`          Scope in {Config, Model, NerfMLP, PropMLP, MLP}
with literals int / float / quoted string / True / False / None / tuples, `#` comments, and the two
references `@jnp.reciprocal` (Model.raydist_fn) and `@coord.contract` ({Nerf,Prop}MLP.warp_fn).
Unknown names raise unless skip_unknown=True (the reference passes skip_unknown=True, configs.py:198).
"""
import ast
import dataclasses
import os
from typing import Any, Dict, List, Optional, Tuple


@dataclasses.dataclass
class Config:
  """The fields of configs.Config the volume-rendering path reads (same names and defaults)."""
  dataset_loader: str = 'llff'
  batch_size: int = 16384
  patch_size: int = 1
  patch_dilation: int = 1
  image_num_per_batch: int = 64
  factor: int = 0
  randomized: bool = True
  near: float = 2.
  far: float = 6.
  checkpoint_dir: Optional[str] = None
  render_dir: Optional[str] = None
  data_dir: Optional[str] = None
  render_chunk_size: int = 16384
  vis_num_rays: int = 16
  transient_type: Optional[str] = None
  max_steps: int = 250000
  early_exit_steps: Optional[int] = None
  checkpoint_every: int = 25000
  print_every: int = 100
  train_render_every: int = 5000
  data_loss_type: str = 'charb'
  charb_padding: float = 0.001
  data_loss_mult: float = 1.0
  data_coarse_loss_mult: float = 0.
  interlevel_loss_mult: float = 1.0
  weight_decay_mults: Dict[str, Any] = dataclasses.field(default_factory=dict)
  disable_multiscale_loss: bool = False
  lr_init: float = 0.002
  lr_final: float = 0.00002
  lr_delay_steps: int = 512
  lr_delay_mult: float = 0.01
  adam_beta1: float = 0.9
  adam_beta2: float = 0.999
  adam_eps: float = 1e-6
  grad_max_norm: float = 0.001
  grad_max_val: float = 0.
  distortion_loss_mult: float = 0.01
  enable_render_zero_glo: bool = False
  enable_render_zero_tra: bool = False
  withmask_transient_weight: float = 0
  static_mask_dir_name: str = 'static_masks'
  finetune_enable: bool = False
  eval_only_once: bool = True
  # engine-side knobs (not in the reference): precision of the MLP path, rays per call
  precision: str = 'bf16_tc'
  extra: Dict[str, Any] = dataclasses.field(default_factory=dict)   # bindings this path does not read


@dataclasses.dataclass
class ModelBindings:
  """gin-bound fields of models.Model (models.py:47-71)."""
  num_prop_samples: int = 64
  num_nerf_samples: int = 32
  num_levels: int = 3
  bg_intensity_range: Tuple[float, float] = (1., 1.)
  anneal_slope: float = 10
  stop_level_grad: bool = True
  use_viewdirs: bool = True
  raydist_fn: Optional[str] = None
  ray_shape: str = 'cone'
  disable_integration: bool = False
  single_jitter: bool = True
  dilation_multiplier: float = 0.5
  dilation_bias: float = 0.0025
  num_glo_features: int = 0
  num_transient_features: int = 0
  num_embeddings: int = 3500
  near_anneal_rate: Optional[float] = None
  near_anneal_init: float = 0.95
  resample_padding: float = 0.0
  use_gpu_resampling: bool = False
  opaque_background: bool = False


@dataclasses.dataclass
class MLPBindings:
  """gin-bound fields of models.MLP / NerfMLP / PropMLP (models.py:360-391)."""
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
  warp_fn: Optional[str] = None
  basis_shape: str = 'icosahedron'
  basis_subdivisions: int = 2


_REFERENCES = {'@jnp.reciprocal': 'reciprocal', '@jnp.log': 'log', '@coord.contract': 'contract',
               '@math.safe_exp': 'safe_exp'}


@dataclasses.dataclass
class Bindings:
  config: Config = dataclasses.field(default_factory=Config)
  model: ModelBindings = dataclasses.field(default_factory=ModelBindings)
  nerf_mlp: MLPBindings = dataclasses.field(default_factory=MLPBindings)
  prop_mlp: MLPBindings = dataclasses.field(default_factory=MLPBindings)


def _parse_value(text: str):
  text = text.strip()
  if text in _REFERENCES:
    return _REFERENCES[text]
  if text.startswith('@'):
    raise ValueError(f'unsupported gin reference {text!r}')
  return ast.literal_eval(text)


def parse_bindings(lines: List[str], bindings: Optional[Bindings] = None, skip_unknown: bool = True) -> Bindings:
  """Applies `Scope.field = value` lines (file contents or --gin_bindings strings) in order."""
  b = bindings or Bindings()
  scopes = {'Config': [b.config], 'Model': [b.model], 'NerfMLP': [b.nerf_mlp], 'PropMLP': [b.prop_mlp],
            'MLP': [b.nerf_mlp, b.prop_mlp]}
  for raw in lines:
    line = raw.split('#', 1)[0].strip() if not ("'" in raw or '"' in raw) else _strip_comment(raw)
    if not line:
      continue
    if '=' not in line:
      raise ValueError(f'cannot parse gin line {raw!r}')
    lhs, rhs = line.split('=', 1)
    lhs = lhs.strip()
    if '/' in lhs:                      # "train/Config.x": scopes are opened by the scripts but unused by the files
      lhs = lhs.rsplit('/', 1)[1]
    if '.' not in lhs:
      raise ValueError(f'cannot parse gin binding {raw!r}')
    scope, field = lhs.rsplit('.', 1)
    value = _parse_value(rhs)
    if scope not in scopes:
      if skip_unknown:
        b.config.extra[lhs] = value
        continue
      raise KeyError(f'unknown configurable {scope!r} in {raw!r}')
    for target in scopes[scope]:
      if hasattr(target, field):
        setattr(target, field, value)
      elif skip_unknown or scope == 'Config':
        b.config.extra[lhs] = value     # e.g. dataset / render-only fields this path never reads
      else:
        raise KeyError(f'{scope} has no field {field!r}')
  return b


def _strip_comment(raw: str) -> str:
  out, quote = [], None
  for ch in raw:
    if quote:
      out.append(ch)
      if ch == quote:
        quote = None
    elif ch in '\'"':
      quote = ch
      out.append(ch)
    elif ch == '#':
      break
    else:
      out.append(ch)
  return ''.join(out).strip()


def load_config(gin_configs: Optional[List[str]] = None, gin_bindings: Optional[List[str]] = None,
                save_config: bool = True) -> Config:
  """configs.load_config (configs.py:195-204): parse files then bindings, optionally dump config.gin.

  Returns the Config; the Model / NerfMLP / PropMLP bindings gin would hold globally ride along as
  `config.bindings` so that `train_utils.setup_model(config, rng)` keeps the reference signature.
  """
  b = Bindings()
  for path in gin_configs or []:
    with open(path) as f:
      parse_bindings(f.read().splitlines(), b)
  parse_bindings(list(gin_bindings or []), b)
  if save_config and b.config.checkpoint_dir:
    os.makedirs(b.config.checkpoint_dir, exist_ok=True)
    with open(os.path.join(b.config.checkpoint_dir, 'config.gin'), 'w') as f:
      f.write(operative_config_str(b))
  b.config.bindings = b
  return b.config


def operative_config_str(b: Bindings) -> str:
  lines = []
  for scope, obj in (('Config', b.config), ('Model', b.model), ('NerfMLP', b.nerf_mlp), ('PropMLP', b.prop_mlp)):
    for f in dataclasses.fields(obj):
      if f.name == 'extra':
        continue
      lines.append(f'{scope}.{f.name} = {getattr(obj, f.name)!r}')
  return '\n'.join(lines) + '\n'
