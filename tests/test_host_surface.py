"""Host-side mirror of the reference call surface (no GPU): gin-subset parser, geopoly basis, Rays/Batch,
render_image chunk / pad / shard logic with a fake render function, lr schedule."""
import glob
import os

import numpy as np
import pytest
import torch

from nerf_hugs_b200.internal import configs, geopoly, math as hmath, utils

G = os.path.join(os.path.dirname(__file__), 'golden')
REF_GINS = '/root/reference/MipNeRF360/configs'


def test_generate_basis_matches_reference_order_exactly():
  """geopoly.generate_basis (geopoly.py:78): same directions in the same ORDER (fixes the IPE column order)."""
  g = np.load(f'{G}/geopoly_basis.npz')
  for key, args in {'icosahedron_2': ('icosahedron', 2), 'octahedron_1': ('octahedron', 1),
                    'octahedron_4': ('octahedron', 4), 'icosahedron_1': ('icosahedron', 1)}.items():
    np.testing.assert_array_equal(geopoly.generate_basis(*args), g[key])
  with pytest.raises(ValueError):
    geopoly.generate_basis('cube', 2)


GIN_360 = """
Config.dataset_loader = 'llff'
Config.near = 0.2
Config.far = 1e6
Config.factor = 4   # a comment

Model.raydist_fn = @jnp.reciprocal
Model.opaque_background = True

PropMLP.warp_fn = @coord.contract
PropMLP.net_depth = 4
PropMLP.net_width = 256
PropMLP.disable_rgb = True

NerfMLP.warp_fn = @coord.contract
NerfMLP.net_depth = 8
NerfMLP.net_width = 1024
"""


def test_gin_subset_parser(tmp_path):
  """configs.load_config (configs.py:195-204) on the grammar of MipNeRF360/configs/360.gin + --gin_bindings."""
  f = tmp_path / '360.gin'
  f.write_text(GIN_360)
  c = configs.load_config([str(f)], ["Config.data_dir = '/data/x # y'", 'Config.batch_size = 4096',
                                     'NerfMLP.net_width = 256', "Config.transient_type = 'withmask'",
                                     'Model.num_glo_features = 48'], save_config=False)
  b = c.bindings
  assert (c.near, c.far, c.factor, c.batch_size, c.data_dir) == (0.2, 1e6, 4, 4096, '/data/x # y')
  assert c.transient_type == 'withmask' and b.model.num_glo_features == 48
  assert b.model.raydist_fn == 'reciprocal' and b.model.opaque_background is True
  assert b.nerf_mlp.warp_fn == 'contract' and b.nerf_mlp.net_width == 256 and b.prop_mlp.disable_rgb is True
  with pytest.raises(ValueError):
    configs.parse_bindings(['Model.raydist_fn = @jnp.square'])
  with pytest.raises(KeyError):
    configs.parse_bindings(['Model.no_such_field = 1'], skip_unknown=False)


@pytest.mark.skipif(not os.path.isdir(REF_GINS), reason='reference tree not present (GPU box)')
def test_all_shipped_gin_files_parse():
  files = sorted(glob.glob(f'{REF_GINS}/*.gin'))
  assert len(files) == 19
  for f in files:
    c = configs.load_config([f], [], save_config=False)
    assert c.bindings.prop_mlp.net_width >= 64


def test_config_dump_roundtrip(tmp_path):
  c = configs.load_config([], [f"Config.checkpoint_dir = '{tmp_path}'", 'Model.num_levels = 2'], save_config=True)
  text = (tmp_path / 'config.gin').read_text()
  c2 = configs.load_config([str(tmp_path / 'config.gin')], [], save_config=False)
  assert c2.bindings.model.num_levels == 2 and 'Model.num_levels = 2' in text


def test_rays_shard_unshard():
  """utils.shard / unshard (utils.py:117-128)."""
  x = torch.arange(24.).reshape(12, 2)
  s = utils.shard(x, 4)
  assert s.shape == (4, 3, 2)
  assert torch.equal(utils.unshard(s), x)
  assert torch.equal(utils.unshard(s, padding=2), x[:-2])
  assert torch.equal(utils.rank_slice(x, 1, 4), x[3:6])
  r = utils.dummy_rays()
  assert r.origins.shape == (1, 3) and r.embed_idx.dtype == torch.int32
  assert set(r.as_dict()) == set(utils.RAY_FIELDS)


def test_render_image_chunking_padding_and_gather():
  """models.render_image (models.py:568-649): chunk loop, edge padding to the device count, unshard,
  per-level ray bundles — with a fake render_fn so the host logic runs without a GPU."""
  from nerf_hugs_b200.internal import models
  H, W, world = 5, 7, 4
  cfg = configs.Config(render_chunk_size=16, vis_num_rays=3)
  rays = utils.Rays(origins=torch.arange(H * W * 3.).reshape(H, W, 3), directions=torch.ones(H, W, 3),
                    viewdirs=torch.ones(H, W, 3), radii=torch.ones(H, W, 1), near=torch.ones(H, W, 1),
                    far=torch.ones(H, W, 1))
  seen = []

  def fake_render(rng, chunk):
    assert chunk.origins.shape[0] == world                       # pre-sharded like the reference's pmap input
    seen.append(chunk.origins.shape[1] * world)
    flat = chunk.origins                                          # [world, n/world, 3]
    out = []
    for level in range(2):
      out.append({'rgb': (flat * (level + 1))[None], 'acc': flat[..., 0][None],
                  'ray_sdist': flat[0][:3, :2][None].expand(1, 3, 2)})
    return out, None

  r = models.render_image(fake_render, rays, None, cfg, verbose=False, world_size=world)
  assert r['rgb'].shape == (H, W, 3) and r['acc'].shape == (H, W)
  assert torch.equal(r['rgb'], rays.origins * 2)                 # last level, padding removed, order preserved
  assert seen == [16, 16, 4]                                      # 35 rays -> chunks 16, 16, 3 (+1 pad)
  assert len(r['ray_sdist']) == 2


def test_learning_rate_decay_matches_oracle():
  from oracle import mipnerf360 as O
  for step in (0, 1, 100, 512, 5000, 250000):
    a = hmath.learning_rate_decay(step, 2e-3, 2e-5, 250000, 512, 0.01)
    b = O.learning_rate_decay(step, 2e-3, 2e-5, 250000, 512, 0.01)
    assert abs(a - b) <= 1e-15


def test_load_dataset_dispatch_is_loud_for_disk_loaders():
  # datasets.load_dataset (datasets.py:45-77): same arguments; file readers are out of scope and say so
  from nerf_hugs_b200.internal import configs, datasets
  config = configs.load_config([], ["Config.dataset_loader = 'phototourism'"], save_config=False)
  with pytest.raises(NotImplementedError, match='from_reference'):
    datasets.load_dataset('train', True, False, 4096, 16, 1, 16, '/tmp/x', config)
  config = configs.load_config([], ["Config.dataset_loader = 'nope'"], save_config=False)
  with pytest.raises(KeyError):
    datasets.load_dataset('train', True, False, 4096, 16, 1, 16, '/tmp/x', config)


def test_image_writers_round_trip(tmp_path):
  # utils.save_img_u8 / save_img_f32 (utils.py:152-163) and the asynchronous writer on host arrays
  import numpy as np
  from PIL import Image
  from nerf_hugs_b200.internal import utils
  rng = np.random.default_rng(0)
  img = rng.uniform(-0.2, 1.2, size=(7, 5, 3)).astype(np.float32)
  img[0, 0, 0] = np.nan
  depth = rng.uniform(0, 9, size=(7, 5)).astype(np.float32)
  wr = utils.AsyncImageWriter(num_workers=1)
  wr.submit_u8(img, str(tmp_path / 'a.png'))
  wr.submit_f32(depth, str(tmp_path / 'd.tiff'))
  assert sorted(os.path.basename(p) for p in wr.close()) == ['a.png', 'd.tiff']
  want = (np.clip(np.nan_to_num(img), 0., 1.) * 255.).astype(np.uint8)
  assert np.array_equal(np.array(Image.open(tmp_path / 'a.png')), want)
  assert np.array_equal(np.array(Image.open(tmp_path / 'd.tiff')), depth)


def test_frame_stripes_cover_the_frame_once():
  # models.render_frame: every rank renders one stripe of rows; the gathered, padded stripes concatenate to the frame
  from nerf_hugs_b200.internal import models
  for h in (1, 7, 8, 9, 800, 1080, 2160):
    for world in (1, 2, 3, 8):
      covered = []
      for rank in range(world):
        rows, r0, r1 = models.frame_stripe(h, rank, world)
        assert 0 <= r0 <= r1 <= h and r1 - r0 <= rows and r0 == min(rank * rows, h)
        covered += list(range(r0, r1))
      assert covered == list(range(h))
      assert rows * world >= h
