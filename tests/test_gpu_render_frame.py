"""Full-frame render pipeline: models.render_frame (one hugs_render_frame call per frame: on-device ray
generation, chunking, frame-sized outputs, uint8 quantisation, on-device squared error) against the reference-shaped route
generate_ray_batch -> models.render_image -> NumPy metrics (eval.py:104-160, utils.py:152-157, image.py mse_to_psnr)."""
import math
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_raygen import _dataset, _device
from tests.test_gpu_surface import _config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('u8,glo', [(False, 0), (True, 4)])
def test_render_frame_equals_render_image(u8, glo, tmp_path):
  from nerf_hugs_b200.internal import models, train_utils, utils
  config = _config(128, glo)
  model, state, render_eval_pfn, _, _ = train_utils.setup_model(config, rng=0, max_rays=300)   # 300: ragged chunks
  ds = _dataset(u8=u8)
  dd = _device(ds)
  cam = 1
  h, w = int(ds['heights'][cam]), int(ds['widths'][cam])
  # reference-shaped route
  batch = dd.generate_ray_batch(cam)
  want = models.render_image(lambda rng, rr: render_eval_pfn(state.params, 0.3, None, rr), batch.rays, None, config,
                             verbose=False)
  got = models.render_frame(model, state.params, dd, cam, 0.3, config, compute_extras=True, want_u8=True, want_psnr=True)
  for k in ('rgb', 'acc', 'distance_mean', 'distance_median'):
    assert got[k].shape == want[k].shape, k
    assert torch.equal(got[k], want[k]), k          # same rays, same kernels: bit-identical
  rgb = want['rgb'].cpu().numpy()
  u8_want = (np.clip(np.nan_to_num(rgb), 0., 1.) * 255.).astype(np.uint8)
  assert np.array_equal(got['rgb_u8'].cpu().numpy(), u8_want)
  gt = np.asarray(ds['images'][cam], np.float64) / (255. if u8 else 1.)
  mse = float(np.mean((rgb.astype(np.float64) - gt) ** 2))
  mse_q = float(np.mean((np.round(rgb.astype(np.float64) * 255) / 255 - gt) ** 2))
  assert abs(got['psnr'] - (-10. / math.log(10.) * math.log(mse))) < 1e-3
  assert abs(got['psnr_quantized'] - (-10. / math.log(10.) * math.log(mse_q))) < 1e-3
  # a stripe of rows equals the rows of the frame
  part = model.engine.render_frame(state.params, dd._cs, cam, w, h, 2, 5, 0.3, zero_glo=config.enable_render_zero_glo)
  assert torch.equal(part['rgb'], want['rgb'][2:5]) and torch.equal(part['acc'], want['acc'][2:5])
  # asynchronous writers: the files hold what utils.save_img_u8 / save_img_f32 of the reference would write
  from PIL import Image
  wr = utils.AsyncImageWriter()
  wr.submit_u8(got['rgb_u8'], str(tmp_path / 'color.png'))
  wr.submit_u8(got['rgb'], str(tmp_path / 'color_f.png'))
  wr.submit_f32(got['distance_mean'], str(tmp_path / 'depth.tiff'))
  assert len(wr.close()) == 3
  assert np.array_equal(np.array(Image.open(tmp_path / 'color.png')), u8_want)
  assert np.array_equal(np.array(Image.open(tmp_path / 'color_f.png')), u8_want)
  assert np.array_equal(np.array(Image.open(tmp_path / 'depth.tiff')), np.nan_to_num(want['distance_mean'].cpu().numpy()))


def test_render_frame_rejects_bad_stripes():
  from nerf_hugs_b200.internal import train_utils
  config = _config(128)
  model, state, *_ = train_utils.setup_model(config, rng=0, max_rays=256)
  dd = _device(_dataset())
  h, w = int(dd.heights_np[0]), int(dd.widths_np[0])
  with pytest.raises(RuntimeError, match='stripe'):
    model.engine.render_frame(state.params, dd._cs, 0, w, h, 3, h + 1, 0.5)
  out = model.engine.render_frame(state.params, dd._cs, 0, w, h, 4, 4, 0.5)      # empty stripe
  assert out['rgb'].shape == (0, w, 3)
