"""`from models import model_config_dict, model_dict, criterion_dict` of the reference scripts
(/root/reference/nerfacto/models/__init__.py:4-17)."""
from . import nerf

model_config_dict = {
    'nerf': nerf.ModelConfig,
}

model_dict = {
    'nerf': nerf.Model,
}

criterion_dict = {
    'nerf': nerf.Loss,
}

try:
  from . import nerfacto
  model_config_dict['nerfacto'] = nerfacto.ModelConfig
  model_dict['nerfacto'] = nerfacto.Model
  criterion_dict['nerfacto'] = nerfacto.Loss
except ImportError:      # pragma: no cover - the hash-grid field is optional while it is being built
  pass
