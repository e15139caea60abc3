"""Rays / Batch containers and shard helpers (MipNeRF360/internal/utils.py:44-126)."""
import dataclasses
from typing import Any, Optional

import torch

RAY_FIELDS = ('pix_coords', 'origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'static_mask', 'near', 'far',
              'embed_idx', 'cam_idx')


@dataclasses.dataclass
class Rays:
  """utils.py:44-57.  All tensors share their leading dims; the last dim is the field width."""
  pix_coords: Any = None
  origins: Any = None
  directions: Any = None
  viewdirs: Any = None
  radii: Any = None
  lossmult: Any = None
  static_mask: Any = None
  near: Any = None
  far: Any = None
  embed_idx: Any = None
  cam_idx: Any = None

  def map(self, fn):
    return Rays(**{k: (None if getattr(self, k) is None else fn(getattr(self, k))) for k in RAY_FIELDS})

  def as_dict(self):
    return {k: getattr(self, k) for k in RAY_FIELDS if getattr(self, k) is not None}


@dataclasses.dataclass
class Batch:
  """utils.py:77-81."""
  rays: Rays
  rgb: Optional[Any] = None


def dummy_rays() -> Rays:
  """utils.py:61-74."""
  z = lambda n: torch.zeros(1, n)
  return Rays(pix_coords=z(2), origins=z(3), directions=z(3), viewdirs=z(3), radii=z(1), lossmult=z(1),
              static_mask=z(1), near=z(1), far=z(1), embed_idx=z(1).int(), cam_idx=z(1).int())


def shard(x, num_shards):
  """utils.py:117-120 with an explicit shard count (the reference uses jax.local_device_count())."""
  return x.reshape((num_shards, -1) + tuple(x.shape[1:]))


def unshard(x, padding=0):
  """utils.py:123-128."""
  y = x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:]))
  return y[:-padding] if padding > 0 else y


def rank_slice(x, rank, world):
  """This rank's contiguous share of a global batch (what utils.shard + pmap give device `rank`)."""
  n = x.shape[0]
  assert n % world == 0, f'global batch {n} is not divisible by the number of ranks {world}'
  per = n // world
  return x[rank * per:(rank + 1) * per]
