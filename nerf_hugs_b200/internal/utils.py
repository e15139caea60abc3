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


# ------------------------------------------------------------------------------------------------------------------
# image writers of eval.py:156-179 / render.py:175-187 (utils.py:152-163)
# ------------------------------------------------------------------------------------------------------------------
def _to_numpy(img):
  import numpy as np
  if torch.is_tensor(img):
    img = img.detach().cpu().numpy()
  return np.asarray(img)


def save_img_u8(img, pth):
  """utils.save_img_u8 (utils.py:152-157): an image in [0, 1] (or already-quantised uint8 bytes) as a uint8 PNG."""
  import numpy as np
  from PIL import Image
  a = _to_numpy(img)
  if a.dtype != np.uint8:
    a = (np.clip(np.nan_to_num(a), 0., 1.) * 255.).astype(np.uint8)
  with open(pth, 'wb') as f:
    Image.fromarray(a).save(f, 'PNG')


def save_img_f32(depthmap, pth):
  """utils.save_img_f32 (utils.py:160-163): a float32 TIFF."""
  import numpy as np
  from PIL import Image
  with open(pth, 'wb') as f:
    Image.fromarray(np.nan_to_num(_to_numpy(depthmap)).astype(np.float32)).save(f, 'TIFF')


class AsyncImageWriter:
  """Writes rendered frames off the render loop: `submit_u8 / submit_f32` copy the device tensor into a pinned host buffer
  on a side stream (the next frame renders meanwhile) and a worker thread encodes the PNG / TIFF.  `close()` drains."""

  def __init__(self, num_workers: int = 2):
    from concurrent.futures import ThreadPoolExecutor
    self._pool = ThreadPoolExecutor(max_workers=num_workers)
    self._futures = []
    self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None

  def _stage(self, img):
    if not (torch.is_tensor(img) and img.is_cuda):
      return img, None
    host = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
    self._stream.wait_stream(torch.cuda.current_stream(img.device))
    with torch.cuda.stream(self._stream):
      host.copy_(img, non_blocking=True)
      ev = torch.cuda.Event()
      ev.record(self._stream)
    img.record_stream(self._stream)
    return host, ev

  def _submit(self, fn, img, pth):
    host, ev = self._stage(img)

    def job():
      if ev is not None:
        ev.synchronize()
      fn(host, pth)
      return pth
    self._futures.append(self._pool.submit(job))

  def submit_u8(self, img, pth):
    self._submit(save_img_u8, img, pth)

  def submit_f32(self, img, pth):
    self._submit(save_img_f32, img, pth)

  def close(self):
    done = [f.result() for f in self._futures]
    self._futures = []
    self._pool.shutdown(wait=True)
    return done
