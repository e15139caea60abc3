"""Edge cases of the torch-twin operators through the C ABI: empty and ragged batches, size limits, loud failures
(the reference's eager torch code handles these shapes; the kernels must too, or say why not)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_sampling_sizes_and_empty_batch():
  from nerf_hugs_b200.nerfacto import ops
  g = torch.Generator().manual_seed(0)
  # 1 ray, 2 rays (less than a warp-block), a non-multiple of the block, many bins -> few samples and the reverse
  for n, nb, ns in ((1, 1, 2), (2, 7, 3), (131, 512, 16), (5, 3, 1024)):
    bins = torch.sort(torch.rand(n, nb + 1, generator=g), -1).values.to(DEV)
    w = torch.rand(n, nb, generator=g).to(DEV)
    out, t = ops.sample_intervals(bins, w, 0.7, 0.01, ns, True, False, (0., 1.), 'piecewise',
                                  torch.full((n, 1), 0.05, device=DEV), torch.full((n, 1), 100.0, device=DEV))
    assert out.shape == (n, ns + 1) and torch.isfinite(out).all() and torch.isfinite(t).all()
    assert (out[:, 1:] >= out[:, :-1]).all() and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert (t[:, 1:] >= t[:, :-1]).all()
  # empty batch: nothing to do, no launch error
  out, _ = ops.sample_intervals(torch.zeros(0, 5, device=DEV), torch.zeros(0, 4, device=DEV), 1., 0., 8, False, True, (0., 1.))
  assert out.shape == (0, 9)
  # num_samples == 1 is refused like stepfun / ray_utils would fail on the reflection of a single centre
  from nerf_hugs_b200._lib import HugsError
  with pytest.raises(HugsError, match='num_samples'):
    ops.sample_intervals(torch.rand(3, 5, device=DEV), torch.rand(3, 4, device=DEV), 1., 0., 1, False, True, (0., 1.))


def test_composite_degenerate_rays():
  from nerf_hugs_b200.nerfacto import ops
  n, S = 7, 33
  g = torch.Generator().manual_seed(1)
  raw = torch.randn(n, S, 4, generator=g)
  raw[0, :, 0] = -1e4            # empty ray: density 0 everywhere -> acc 0, depth 0 / eps clipped to 0
  raw[1, :, 0] = 80.0            # opaque at the first sample (softplus passes x > 20 through)
  raw[2, 5, 0] = float('nan')    # a NaN density: nan_to_num on the weights (quirk B6)
  bins = torch.sort(torch.rand(n, S + 1, generator=g) * 3 + 1, -1).values
  bins[3] = 2.0                  # zero-length ray
  dirs = torch.randn(n, 3, generator=g)
  bg = torch.rand(n, 3, generator=g)
  cfg = ops.render_cfg(False, 'softplus', -1.0, 1.0, 0.0, 0.001)
  w, rgb, depth, acc, smax = ops.composite_forward(cfg, raw.to(DEV), bins.to(DEV), dirs.to(DEV), bg.to(DEV))
  assert torch.isfinite(rgb[[0, 1, 3, 4, 5, 6]]).all() and torch.isfinite(w[[0, 1, 3, 4, 5, 6]]).all()
  assert float(acc[0]) == 0.0 and torch.allclose(rgb[0].cpu(), bg[0])          # background only
  assert float(acc[1]) > 0.999                                                  # saturated
  assert float(acc[3]) == 0.0                                                   # no extent, no weight
  assert float(w[2, 5]) == 0.0 and torch.isfinite(w[2, :5]).all()              # the NaN sample itself is zeroed
  d_raw = ops.composite_backward(cfg, raw.to(DEV), bins.to(DEV), dirs.to(DEV), bg.to(DEV), smax,
                                 torch.randn(n, S, device=DEV), torch.randn(n, 3, device=DEV), torch.randn(n, device=DEV),
                                 torch.randn(n, device=DEV))
  assert torch.isfinite(d_raw[[0, 1, 3, 4, 5, 6]]).all()
  # size limits are loud
  from nerf_hugs_b200._lib import HugsError
  with pytest.raises(HugsError, match='samples per ray'):
    ops.composite_forward(cfg, torch.zeros(2, 1, 4, device=DEV), torch.zeros(2, 2, device=DEV), dirs[:2].to(DEV), None)


def test_field_engine_limits_are_loud():
  from nerf_hugs_b200 import _lib
  from nerf_hugs_b200._lib import HugsError, lib, check
  d = _lib.HashFieldDesc()
  d.n_levels, d.features_per_level, d.log2_hashmap_size, d.base_res, d.per_level_scale = 16, 2, 19, 16, 1.38
  d.hidden_dim, d.geo_feat_dim, d.hidden_dim_color, d.bound, d.max_samples, d.max_rays = 64, 15, 64, 1.0, 1024, 64
  h = C.c_void_p()
  rc = lib.hugs_hashfield_create(C.byref(d), C.byref(h))            # nerfacto's dataclass defaults (64 / 15 / 64): not built
  assert rc == -1 and b'256 / 256 / 64' in lib.hugs_last_error()
  d.features_per_level = 4
  assert lib.hugs_hashfield_create(C.byref(d), C.byref(h)) == -1 and b'features_per_level' in lib.hugs_last_error()
  # a density field refuses more samples than its workspace
  d.features_per_level, d.geo_feat_dim, d.hidden_dim_color, d.n_levels = 2, 0, 0, 5
  check(lib.hugs_hashfield_create(C.byref(d), C.byref(h)))
  grid = torch.zeros(int(lib.hugs_hashfield_grid_floats(h)), device=DEV)
  mlp = torch.zeros(int(lib.hugs_hashfield_mlp_floats(h)), device=DEV)
  r = _lib.Rays()
  o = torch.zeros(64, 3, device=DEV); r.origins = o.data_ptr(); r.directions = o.data_ptr()
  td = torch.zeros(64, 33, device=DEV); raw = torch.zeros(64, 32, device=DEV)
  rc = lib.hugs_hashfield_forward(h, grid.data_ptr(), mlp.data_ptr(), C.byref(r), td.data_ptr(), 64, 32, 0, 0, raw.data_ptr(), None)
  assert rc == -1 and b'exceed max_samples' in lib.hugs_last_error()
  # all-zero parameters on a degenerate ray: finite output (bias 0 -> raw density 0), nothing launched out of bounds
  check(lib.hugs_hashfield_forward(h, grid.data_ptr(), mlp.data_ptr(), C.byref(r), td.data_ptr(), 32, 32, 0, 0, raw.data_ptr(), None))
  torch.cuda.synchronize()
  assert torch.isfinite(raw[:32]).all()
  lib.hugs_hashfield_destroy(h)


def test_merge_bins_single_interval_inputs():
  from nerf_hugs_b200.nerfacto import ops
  a = torch.tensor([[0.0, 1.0], [0.2, 0.4]], device=DEV)           # one coarse interval
  b = torch.tensor([[0.0, 0.5, 1.0], [0.2, 0.2, 0.4]], device=DEV)
  out, _ = ops.merge_bins(a, b, (0., 1.))
  assert out.shape == (2, 4) and (out[:, 1:] >= out[:, :-1]).all()
  assert float(out[0, 0]) >= 0.0 and float(out[0, -1]) <= 1.0
