"""Device-resident dataset and batch assembly (MipNeRF360/internal/datasets.py:259-529, camera_utils.py:503-669).

The reference builds every training batch on the host (NumPy gathers + pixels_to_rays in a producer thread,
`datasets.py:393-407`) and ships 84 B per ray to the device.  Here the images, HuGS static masks, near/far maps and
cameras live in HBM once, and a batch is produced by one kernel (`hugs_make_ray_batch`) from integer
(camera, x, y) draws made on the device — same fields, shapes and value semantics as `Dataset._make_ray_batch`.

Lens distortion (`Dataset.distortion_params`, one set per dataset as in the reference) and fisheye cameras are handled in
the kernel; NDC and the spherical render paths are out of scope (NotImplementedError).
"""
import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib
from . import utils


def _ptr(t):
  return None if t is None else t.data_ptr()


class DeviceDataset:
  """Holds Dataset.{images, static_masks, nears, fars, pixtocams, camtoworlds, heights, widths, embed_idxs} on the device.

  Args mirror the attributes `datasets.Dataset` fills in `_load_renderings` (datasets.py:310-383):
    images: list of [H_i, W_i, 3] arrays (float in [0, 1] or uint8), or None for render-only sets;
    static_masks / nears / fars: lists of [H_i, W_i, 1] (or [H_i, W_i]) float arrays, or None;
    pixtocams [N, 3, 3], camtoworlds [N, 3, 4]; embed_idxs [N] or None (camera index).
  """

  def __init__(self, pixtocams, camtoworlds, heights, widths, images: Optional[Sequence] = None,
               static_masks: Optional[Sequence] = None, nears: Optional[Sequence] = None,
               fars: Optional[Sequence] = None, embed_idxs=None, near: float = 0.2, far: float = 1e6,
               distortion_params=None, camtype: str = 'perspective', device=None):
    if camtype not in ('perspective', 'fisheye'):
      raise NotImplementedError(f'camera type {camtype!r} is not supported on the device path')
    if distortion_params is not None and not isinstance(distortion_params, dict):
      raise NotImplementedError('distortion_params must be one dict of k1..k4, p1, p2 for the whole dataset (as '
                                'datasets.Dataset.distortion_params); per-camera lists are not supported')
    if not torch.cuda.is_available():
      raise RuntimeError('DeviceDataset needs a CUDA device (the product path has no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self.device = dev
    self.heights_np = np.asarray(heights, np.int64); self.widths_np = np.asarray(widths, np.int64)
    self.n_cams = len(self.heights_np)
    sizes = self.heights_np * self.widths_np
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    f32 = lambda a: torch.as_tensor(np.array(a, np.float32, copy=True), device=dev)
    self.pixtocams = f32(np.broadcast_to(np.asarray(pixtocams, np.float32), (self.n_cams, 3, 3)))
    self.camtoworlds = f32(np.asarray(camtoworlds)[..., :3, :4])
    self.heights = torch.as_tensor(self.heights_np.astype(np.int32), device=dev)
    self.widths = torch.as_tensor(self.widths_np.astype(np.int32), device=dev)
    self.pixel_offset = torch.as_tensor(offs.astype(np.int64), device=dev)
    self.embed_idxs = None if embed_idxs is None else torch.as_tensor(np.asarray(embed_idxs, np.int32), device=dev)

    def pack(store, ch):
      if store is None:
        return None
      flat = [np.asarray(a).reshape(-1, ch) for a in store]
      for a, n in zip(flat, sizes):
        assert a.shape[0] == n, 'per-pixel store does not match heights x widths'
      return np.concatenate(flat, 0)
    self.images = self.images_u8 = None
    if images is not None:
      img = pack(images, 3)
      if img.dtype == np.uint8:
        self.images_u8 = torch.as_tensor(np.ascontiguousarray(img), device=dev)
      else:
        self.images = f32(img)
    sm, nn, ff = pack(static_masks, 1), pack(nears, 1), pack(fars, 1)
    self.static_masks = None if sm is None else f32(sm[:, 0])
    self.nears = None if nn is None else f32(nn[:, 0])
    self.fars = None if ff is None else f32(ff[:, 0])
    self.near, self.far = float(near), float(far)
    cs = _lib.CameraSet()
    cs.pixtocams, cs.camtoworlds = _ptr(self.pixtocams), _ptr(self.camtoworlds)
    cs.heights, cs.widths, cs.pixel_offset = _ptr(self.heights), _ptr(self.widths), _ptr(self.pixel_offset)
    cs.images, cs.images_u8 = _ptr(self.images), _ptr(self.images_u8)
    cs.static_masks, cs.nears, cs.fars = _ptr(self.static_masks), _ptr(self.nears), _ptr(self.fars)
    cs.embed_idxs = _ptr(self.embed_idxs)
    cs.near, cs.far = self.near, self.far
    self.distortion = None
    if distortion_params is not None:
      unknown = set(distortion_params) - {'k1', 'k2', 'k3', 'k4', 'p1', 'p2'}
      if unknown:
        raise ValueError(f'unknown distortion parameters {sorted(unknown)}')
      self.distortion = f32([distortion_params.get(k, 0.0) for k in ('k1', 'k2', 'k3', 'k4', 'p1', 'p2')])
    cs.distortion = _ptr(self.distortion)
    cs.camtype = {'perspective': 0, 'fisheye': 1}[camtype]
    self._cs = cs

  # datasets.py:446-482
  def make_ray_batch(self, pix_x_int, pix_y_int, cam_idx, want_rgb: bool = True) -> utils.Batch:
    """Rays + ground truth of pixels (cam_idx, pix_y_int, pix_x_int): int tensors of one common shape."""
    dev = self.device
    i32 = lambda t: torch.as_tensor(t, device=dev).to(torch.int32).contiguous()
    px, py = i32(pix_x_int), i32(pix_y_int)
    ci = i32(cam_idx).expand_as(px).contiguous() if torch.as_tensor(cam_idx).ndim == 0 else i32(cam_idx)
    shape = tuple(px.shape)
    n = px.numel()
    mk = lambda c, dt=torch.float32: torch.empty(shape + (c,), device=dev, dtype=dt)
    rays = utils.Rays(pix_coords=mk(2), origins=mk(3), directions=mk(3), viewdirs=mk(3), radii=mk(1), lossmult=mk(1),
                      static_mask=mk(1), near=mk(1), far=mk(1), embed_idx=mk(1, torch.int32), cam_idx=mk(1, torch.int32))
    has_img = self.images is not None or self.images_u8 is not None
    rgb = mk(3) if (want_rgb and has_img) else None
    out = _lib.RayBatch()
    for k in ('origins', 'directions', 'viewdirs', 'radii', 'near', 'far', 'lossmult', 'static_mask', 'embed_idx',
              'cam_idx', 'pix_coords'):
      setattr(out, k, getattr(rays, k).data_ptr())
    out.rgb = _ptr(rgb)
    if n > 0:
      with torch.cuda.device(dev):
        check(lib.hugs_make_ray_batch(C.byref(self._cs), ci.data_ptr(), px.data_ptr(), py.data_ptr(), n, C.byref(out),
                                      torch.cuda.current_stream(dev).cuda_stream))
    return utils.Batch(rays=rays, rgb=rgb)

  # datasets.py:484-527 (_next_train): random patches of `image_num_per_batch` random cameras
  def next_train_batch(self, gen: torch.Generator, batch_size: int, patch_size: int = 1, patch_dilation: int = 1,
                       image_num_per_batch: int = 1, sample_from_half_image: bool = False) -> utils.Batch:
    dev = self.device
    n_patch = (batch_size // image_num_per_batch) // patch_size ** 2
    upper = (patch_size - 1) * patch_dilation
    cams = torch.randint(0, self.n_cams, (image_num_per_batch,), generator=gen, device=dev)
    h = self.heights[cams].to(torch.float32); w = self.widths[cams].to(torch.float32)
    if sample_from_half_image:
      w = torch.floor(w / 2)
    u = torch.rand(2, image_num_per_batch, n_patch, generator=gen, device=dev)
    x0 = torch.floor(u[0] * (w[:, None] - upper)).to(torch.int32)
    y0 = torch.floor(u[1] * (h[:, None] - upper)).to(torch.int32)
    d = torch.arange(patch_size, device=dev, dtype=torch.int32) * patch_dilation
    px = (x0[:, :, None, None] + d[None, None, None, :]).expand(-1, -1, patch_size, -1)     # pixel_coordinates: x varies fastest
    py = (y0[:, :, None, None] + d[None, None, :, None]).expand(-1, -1, -1, patch_size)
    ci = cams.to(torch.int32)[:, None, None, None].expand_as(px)
    flat = lambda t: t.reshape(image_num_per_batch * n_patch, patch_size, patch_size).contiguous()
    return self.make_ray_batch(flat(px), flat(py), flat(ci))

  # datasets.py:529-560 (generate_ray_batch): every pixel of one camera
  def generate_ray_batch(self, cam_idx: int) -> utils.Batch:
    h, w = int(self.heights_np[cam_idx]), int(self.widths_np[cam_idx])
    ys, xs = torch.meshgrid(torch.arange(h, device=self.device), torch.arange(w, device=self.device), indexing='ij')
    return self.make_ray_batch(xs, ys, torch.full_like(xs, cam_idx))


# ------------------------------------------------------------------------------------------------------------------
# datasets.load_dataset / datasets.Dataset: the iterator the scripts consume (datasets.py:45-77, 259-443)
# ------------------------------------------------------------------------------------------------------------------
class Dataset:
  """Iterator protocol of `datasets.Dataset` (datasets.py:259-443) over a DeviceDataset.

  The reference fills a `queue.Queue(3)` from a daemon thread that assembles batches with NumPy (`:393-407`); here a batch is
  one kernel launch on the device, so `__next__` produces it on demand - same `utils.Batch`, same `size`, `peek`,
  `generate_ray_batch`.  Training batches are this rank's share (`batch_size // world_size` rays: the reference shards with
  `utils.shard`, utils.py:117-120); test examples are whole images in order.

  Subclasses implement `_load_renderings(config)` exactly as in the reference: set `images` (list of [H, W, 3]),
  `camtoworlds`, `pixtocams`, `heights`, `widths` and optionally `static_masks`, `nears`, `fars`, `embed_idxs`,
  `distortion_params`, `camtype`.  Reading COLMAP / Blender / Phototourism files is out of this package's scope (SURVEY.md
  §2.1): `from_reference(ds)` adopts the arrays of a dataset object loaded by the reference's own loaders instead.
  """

  def __init__(self, split, is_training, sample_from_half_image, batch_size, patch_size, patch_dilation, image_num_per_batch,
               data_dir, config, rank=0, world_size=1, seed=20221019):
    self.split, self.is_training = split, bool(is_training)
    self.sample_from_half_image = sample_from_half_image
    self._world = max(1, int(world_size))
    self._batch_size = batch_size // self._world
    self._patch_size, self._patch_dilation = max(patch_size, 1), patch_dilation
    self._image_num_per_batch = max(1, image_num_per_batch // self._world) if image_num_per_batch >= self._world else 1
    self.data_dir, self.config = data_dir, config
    self.near, self.far = config.near, config.far
    self.static_masks = self.nears = self.fars = self.embed_idxs = self.distortion_params = None
    self.camtype = 'perspective'
    self._load_renderings(config)
    self._n_examples = len(self.camtoworlds)
    self.device_dataset = DeviceDataset(
        self.pixtocams, self.camtoworlds, self.heights, self.widths, images=self.images, static_masks=self.static_masks,
        nears=self.nears, fars=self.fars, embed_idxs=self.embed_idxs, near=self.near, far=self.far,
        distortion_params=self.distortion_params, camtype=self.camtype)
    self._gen = torch.Generator(device=self.device_dataset.device)
    self._gen.manual_seed(seed + rank)
    self._test_idx = 0
    self._peeked = None

  def _load_renderings(self, config):
    raise NotImplementedError

  @classmethod
  def from_reference(cls, ds, is_training, sample_from_half_image, batch_size, patch_size, patch_dilation,
                     image_num_per_batch, config, **kw):
    """Adopt the arrays of a `datasets.Dataset` loaded by the reference's own loaders (attributes of datasets.py:310-383)."""
    class _Adopted(cls):
      def _load_renderings(self, config):
        for k in ('images', 'camtoworlds', 'pixtocams', 'static_masks', 'nears', 'fars', 'embed_idxs', 'distortion_params'):
          setattr(self, k, getattr(ds, k, None))
        n = len(ds.camtoworlds)
        hs, ws = getattr(ds, 'heights', None), getattr(ds, 'widths', None)
        self.heights = hs if hs is not None else [ds.height] * n
        self.widths = ws if ws is not None else [ds.width] * n
    return _Adopted(ds.split, is_training, sample_from_half_image, batch_size, patch_size, patch_dilation,
                    image_num_per_batch, getattr(ds, 'data_dir', None), config, **kw)

  def __iter__(self):
    return self

  def _make(self):
    if self.is_training:
      return self.device_dataset.next_train_batch(self._gen, self._batch_size, self._patch_size, self._patch_dilation,
                                                  self._image_num_per_batch, self.sample_from_half_image)
    idx = self._test_idx
    self._test_idx = (self._test_idx + 1) % self._n_examples
    return self.device_dataset.generate_ray_batch(idx)

  def __next__(self):
    if self._peeked is not None:
      b, self._peeked = self._peeked, None
      return b
    return self._make()

  def peek(self):
    if self._peeked is None:
      self._peeked = self._make()
    return self._peeked

  @property
  def size(self):
    return self._n_examples

  def generate_ray_batch(self, cam_idx: int) -> utils.Batch:
    return self.device_dataset.generate_ray_batch(cam_idx)


class Synthetic(Dataset):
  """The synthetic scene of the measurement configs: cameras on a unit sphere looking at the origin, uniform
  random images; with `config.transient_type == 'withmask'` also 32 x 32-block Bernoulli(0.8) HuGS static masks and per-image
  near / far (config 3's Phototourism shape).  `Config.dataset_loader = 'synthetic'`."""

  n_cams, hw = 16, (200, 200)

  def _load_renderings(self, config):
    rng = np.random.default_rng(0 if self.split == 'train' else 1)
    n, (h, w) = self.n_cams, self.hw
    pos = rng.normal(size=(n, 3)); pos /= np.linalg.norm(pos, axis=-1, keepdims=True)
    fwd = -pos
    right = np.cross(fwd, np.array([0., 0., 1.])); right /= np.linalg.norm(right, axis=-1, keepdims=True) + 1e-9
    up = np.cross(right, fwd)
    self.camtoworlds = np.stack([right, up, -fwd, pos], -1).astype(np.float32)
    focal = 1111.1 / 800. * w
    self.pixtocams = np.linalg.inv(np.array([[focal, 0, w / 2.], [0, focal, h / 2.], [0, 0, 1.]])).astype(np.float32)
    self.heights, self.widths = [h] * n, [w] * n
    self.images = [rng.uniform(size=(h, w, 3)).astype(np.float32) for _ in range(n)]
    if getattr(config, 'transient_type', None) == 'withmask':
      blocks = lambda: (rng.uniform(size=((h + 31) // 32, (w + 31) // 32)) < 0.8).astype(np.float32)
      self.static_masks = [np.kron(blocks(), np.ones((32, 32), np.float32))[:h, :w, None] for _ in range(n)]
    self.embed_idxs = np.arange(n)


_LOADERS = {'synthetic': Synthetic}
_REFERENCE_LOADERS = ('blender', 'llff', 'tat_nerfpp', 'tat_fvs', 'dtu', 'kubric', 'phototourism', 'distractor')


def load_dataset(split, is_training, sample_from_half_image, batch_size, patch_size, patch_dilation, image_num_per_batch,
                 train_dir, config, **kw):
  """datasets.load_dataset (datasets.py:45-77): the same nine arguments, dispatched on `config.dataset_loader`."""
  name = config.dataset_loader
  if name in _REFERENCE_LOADERS:
    raise NotImplementedError(
        f"dataset_loader={name!r}: reading {name} files from disk (images, COLMAP / json poses, HuGS mask files) is outside this "
        "package's scope (SURVEY.md §2.1).  Load the split with the reference's own datasets.load_dataset and hand it over: "
        "nerf_hugs_b200.internal.datasets.Dataset.from_reference(ds, sample_from_half_image=..., batch_size=..., ...)")
  if name not in _LOADERS:
    raise KeyError(name)
  return _LOADERS[name](split, is_training, sample_from_half_image, batch_size, patch_size, patch_dilation,
                        image_num_per_batch, train_dir, config, **kw)
