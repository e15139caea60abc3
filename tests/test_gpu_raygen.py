"""hugs_make_ray_batch (on-device pixels_to_rays + HuGS mask / near / far / colour gather) vs the CPU oracle and the
reference-generated golden rays."""
import os

import numpy as np
import pytest
import torch

from oracle import camera as OC

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camera.npz'))


def _dataset(u8=False, seed=0):
  rng = np.random.default_rng(seed)
  hs, ws = G['heights'], G['widths']
  if u8:
    images = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for h, w in zip(hs, ws)]
  else:
    images = [rng.uniform(size=(h, w, 3)).astype(np.float32) for h, w in zip(hs, ws)]
  return dict(pixtocams=G['pixtocams'], camtoworlds=G['camtoworlds'], heights=hs, widths=ws, images=images,
              static_masks=[(rng.uniform(size=(h, w, 1)) < 0.8).astype(np.float32) for h, w in zip(hs, ws)],
              nears=[rng.uniform(0.1, 0.3, size=(h, w, 1)).astype(np.float32) for h, w in zip(hs, ws)],
              fars=[rng.uniform(2., 9., size=(h, w, 1)).astype(np.float32) for h, w in zip(hs, ws)],
              embed_idxs=np.arange(len(hs))[::-1].copy())


def _device(ds):
  from nerf_hugs_b200.internal.datasets import DeviceDataset
  return DeviceDataset(ds['pixtocams'], ds['camtoworlds'], ds['heights'], ds['widths'], images=ds['images'],
                       static_masks=ds['static_masks'], nears=ds['nears'], fars=ds['fars'], embed_idxs=ds['embed_idxs'])


@pytest.mark.parametrize('u8', [False, True])
def test_make_ray_batch_vs_oracle_and_reference_golden(u8):
  ds = _dataset(u8)
  dd = _device(ds)
  ci, px, py = G['cam_idx'], G['pix_x'], G['pix_y']
  b = dd.make_ray_batch(torch.tensor(px), torch.tensor(py), torch.tensor(ci))
  ref_ds = dict(ds)
  if u8:
    ref_ds['images'] = [im.astype(np.float32) / 255. for im in ds['images']]      # datasets.py image loading
  rays, rgb = OC.make_ray_batch(ref_ds, ci, px, py)
  got = {k: v.cpu().numpy() for k, v in b.rays.as_dict().items()}
  # gathers and integer fields: bit-exact
  for k in ('static_mask', 'near', 'far', 'lossmult', 'embed_idx', 'cam_idx', 'origins'):
    assert np.array_equal(got[k], rays[k]), k
  assert np.array_equal(b.rgb.cpu().numpy(), rgb)
  # geometry: float64 arithmetic rounded to float32 on both sides; fused multiply-adds may flip the last bit
  for k, tol in (('directions', 2e-7), ('viewdirs', 2e-7), ('radii', 2e-7), ('pix_coords', 2e-7)):
    np.testing.assert_allclose(got[k], rays[k], rtol=tol, atol=1e-9, err_msg=k)
    assert (got[k] == rays[k]).mean() > 0.99, k
  # and directly against the reference's own outputs
  np.testing.assert_allclose(got['directions'], G['directions'].astype(np.float32), rtol=2e-7, atol=1e-9)
  np.testing.assert_allclose(got['radii'], G['radii'].astype(np.float32), rtol=2e-7)


def test_generate_ray_batch_and_patch_sampling():
  ds = _dataset()
  dd = _device(ds)
  cam = 2
  b = dd.generate_ray_batch(cam)
  h, w = int(ds['heights'][cam]), int(ds['widths'][cam])
  assert tuple(b.rays.origins.shape) == (h, w, 3) and tuple(b.rgb.shape) == (h, w, 3)
  assert np.array_equal(b.rgb.cpu().numpy(), ds['images'][cam])
  assert np.array_equal(b.rays.static_mask.cpu().numpy(), ds['static_masks'][cam])
  gen = torch.Generator(device='cuda'); gen.manual_seed(0)
  tb = dd.next_train_batch(gen, batch_size=512, patch_size=4, patch_dilation=2, image_num_per_batch=2)
  assert tuple(tb.rays.origins.shape) == (32, 4, 4, 3)           # [patches, patch, patch, 3] as datasets.py:484-527
  pc = tb.rays.pix_coords.cpu().numpy(); ci = tb.rays.cam_idx.cpu().numpy()[..., 0]
  assert ((pc > 0) & (pc < 1)).all()
  x = pc[..., 0] * ds['widths'][ci] - 0.5
  assert np.allclose(x[:, 0, 1] - x[:, 0, 0], 2.0, atol=1e-3)     # dilation along x inside a patch
  assert len(np.unique(ci)) <= 2 and (ci[0] == ci[0, 0, 0]).all()


@pytest.mark.parametrize('tag', ['dist', 'fish', 'distfish'])
def test_lens_distortion_and_fisheye_vs_reference_golden(tag):
  """Newton lens undistortion (camera_utils.py:460-494) and the fisheye projection (:557-568) in the ray-generation
  kernel, against outputs of the reference's own pixels_to_rays."""
  from nerf_hugs_b200.internal.datasets import DeviceDataset
  ds = _dataset()
  dist = dict(zip(('k1', 'k2', 'k3', 'k4', 'p1', 'p2'), [float(v) for v in G['dist_params']]))
  dd = DeviceDataset(ds['pixtocams'], ds['camtoworlds'], ds['heights'], ds['widths'],
                     distortion_params=dist if 'dist' in tag else None, camtype='fisheye' if 'fish' in tag else 'perspective')
  b = dd.make_ray_batch(torch.tensor(G['pix_x']), torch.tensor(G['pix_y']), torch.tensor(G['cam_idx']), want_rgb=False)
  got = {k: v.cpu().numpy() for k, v in b.rays.as_dict().items()}
  for k in ('directions', 'viewdirs', 'radii'):
    np.testing.assert_allclose(got[k], G[f'{tag}_{k}'].astype(np.float32), rtol=3e-7, atol=1e-9, err_msg=k)


def test_unsupported_camera_models_are_loud():
  from nerf_hugs_b200.internal.datasets import DeviceDataset
  ds = _dataset()
  with pytest.raises(NotImplementedError):
    DeviceDataset(ds['pixtocams'], ds['camtoworlds'], ds['heights'], ds['widths'], camtype='pano')
  with pytest.raises(NotImplementedError):
    DeviceDataset(ds['pixtocams'], ds['camtoworlds'], ds['heights'], ds['widths'], distortion_params=[{'k1': 0.1}])


def test_device_batches_drive_the_training_step():
  """A batch assembled on the device (hugs_make_ray_batch) trains exactly like the same batch handed over from the
  host: the two train_pstep calls start from identical states and must report identical statistics."""
  import copy
  from nerf_hugs_b200.internal import configs, train_utils, utils
  ds = _dataset(u8=True)
  dd = _device(ds)
  bind = ['Config.batch_size = 256', 'Config.patch_size = 8', "Config.transient_type = 'withmask'",
          'Model.opaque_background = True', 'Model.num_levels = 2', 'Model.num_prop_samples = 64',
          'Model.num_nerf_samples = 128', 'Model.num_glo_features = 4', 'Model.num_embeddings = 16',
          'PropMLP.net_depth = 4', 'PropMLP.disable_rgb = True', 'NerfMLP.net_width = 256']
  config = configs.load_config([], bind, save_config=False)
  stats = []
  for mode in ('device', 'host'):
    model, state, _, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=256)
    gen = torch.Generator(device='cuda'); gen.manual_seed(3)
    batch = dd.next_train_batch(gen, 256, patch_size=8, patch_dilation=1, image_num_per_batch=4)
    assert tuple(batch.rays.origins.shape) == (4, 8, 8, 3)
    if mode == 'host':
      batch = utils.Batch(rays=batch.rays.map(lambda t: t.cpu()), rgb=batch.rgb.cpu())
    g2 = torch.Generator(device='cuda'); g2.manual_seed(5)
    state, st, _ = train_pstep(g2, state, batch, 0.1, None)
    stats.append((st['loss'], st['losses']['data'], float(state.params.double().sum())))
    assert 0.0 < float(batch.rays.static_mask.float().mean()) < 1.0      # the HuGS masks are really in the batch
  # the forward is deterministic (identical losses); the weight gradients are reduced with fp32 atomics, so the
  # updated parameters agree to rounding only
  assert stats[0][:2] == stats[1][:2], stats
  assert abs(stats[0][2] - stats[1][2]) <= 1e-6 * abs(stats[0][2]) + 1e-6, stats


def test_load_dataset_iterator_protocol():
  # datasets.load_dataset + Dataset.__iter__/__next__/peek/size/generate_ray_batch (datasets.py:45-77, 393-443)
  from nerf_hugs_b200.internal import configs, datasets
  bind = ["Config.dataset_loader = 'synthetic'", "Config.transient_type = 'withmask'", 'Config.near = 0.5', 'Config.far = 3.0']
  config = configs.load_config([], bind, save_config=False)
  ds = datasets.load_dataset('train', True, False, 1024, 16, 1, 2, None, config)
  assert ds.size == 16
  first = ds.peek()
  b = next(iter(ds))
  assert b is first
  assert tuple(b.rgb.shape) == (4, 16, 16, 3) and tuple(b.rays.origins.shape) == (4, 16, 16, 3)     # [patches, 16, 16, C]
  m = b.rays.static_mask
  assert torch.all((m == 0) | (m == 1)) and 0.3 < float(m.mean()) <= 1.0
  assert torch.allclose(b.rays.near, torch.full_like(b.rays.near, 0.5))
  b2 = next(ds)
  assert not torch.equal(b2.rays.pix_coords, b.rays.pix_coords)
  # two ranks of a 2-GPU job draw different, half-sized batches (utils.shard, utils.py:117-120)
  r0 = datasets.load_dataset('train', True, False, 1024, 16, 1, 2, None, config, rank=0, world_size=2)
  r1 = datasets.load_dataset('train', True, False, 1024, 16, 1, 2, None, config, rank=1, world_size=2)
  a0, a1 = next(r0), next(r1)
  assert a0.rgb.shape[0] == 2 and not torch.equal(a0.rays.pix_coords, a1.rays.pix_coords)
  test = datasets.load_dataset('test', False, False, 1024, 16, 1, 2, None, config)
  img = next(test)
  assert tuple(img.rgb.shape) == (200, 200, 3) and int(img.rays.cam_idx.max()) == 0
  assert int(next(test).rays.cam_idx.max()) == 1
