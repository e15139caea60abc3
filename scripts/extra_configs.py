"""Measurements of the BASELINE.json parity configs that are not the bench line (JSON lines on stdout; one GPU, or N GPUs
of one box under `python -m torch.distributed.run --nproc-per-node N scripts/extra_configs.py ...`: rays shard over the
ranks, times are the maximum over ranks, rank 0 prints):

  config 3  Mip-NeRF 360 with HuGS static masks, Phototourism-shape synthetic scene
            (phototourism_1024_withmask.gin shape: no warp_fn / raydist_fn, per-pixel near/far, patch 16, 48 GLO
            features, transient_type='withmask', charb loss, distortion 0.001), batches assembled ON THE DEVICE by
            hugs_make_ray_batch from uint8 images + static masks (DeviceDataset.next_train_batch)
  config 5  full-frame render sweep through models.render_image (deterministic path, compute_extras=True)

Usage: python scripts/extra_configs.py [hugs] [render] [--steps K]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import torch.distributed as dist

from nerf_hugs_b200.internal import configs, models, train_utils, utils
from nerf_hugs_b200.internal.datasets import DeviceDataset

RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))


def device():
  torch.cuda.set_device(LOCAL)
  dev = torch.device('cuda', LOCAL)
  if WORLD > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  return dev


def timed(fn, steps, dev):
  """ms per step: barrier + synchronize on both sides, CUDA events, maximum over ranks."""
  if WORLD > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = None
  for _ in range(steps):
    out = fn()
  e1.record()
  if WORLD > 1:
    dist.barrier()
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
  if WORLD > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return float(ms.item()), out


def emit(d):
  if RANK == 0:
    print(json.dumps(d), flush=True)


def sphere_cameras(rng, n_cams, hw):
  pos = rng.normal(size=(n_cams, 3)); pos /= np.linalg.norm(pos, axis=-1, keepdims=True)
  fwd = -pos
  right = np.cross(fwd, np.array([0., 0., 1.])); right /= np.linalg.norm(right, axis=-1, keepdims=True) + 1e-9
  up = np.cross(right, fwd)
  c2w = np.stack([right, up, -fwd, pos], -1).astype(np.float32)                 # OpenGL: camera looks along -z
  focal = 1111.1 / 800. * hw[:, 1]
  p2c = np.stack([np.linalg.inv(np.array([[f, 0, w / 2.], [0, f, h / 2.], [0, 0, 1.]])) for f, (h, w) in zip(focal, hw)])
  return p2c.astype(np.float32), c2w


def hugs_dataset(n_cams=64, seed=0):
  rng = np.random.default_rng(seed)
  hw = rng.integers(400, 801, size=(n_cams, 2))
  p2c, c2w = sphere_cameras(rng, n_cams, hw)
  images, masks, nears, fars = [], [], [], []
  for h, w in hw:
    images.append(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
    blocks = (rng.uniform(size=((h + 31) // 32, (w + 31) // 32)) < 0.8).astype(np.float32)     # 32x32-block Bernoulli(0.8)
    masks.append(np.kron(blocks, np.ones((32, 32), np.float32))[:h, :w, None])
    nears.append(np.full((h, w, 1), 1.0 * rng.uniform(0.9, 1.1), np.float32))
    fars.append(np.full((h, w, 1), 2.0 * rng.uniform(0.9, 1.1), np.float32))
  return DeviceDataset(p2c, c2w, hw[:, 0], hw[:, 1], images=images, static_masks=masks, nears=nears, fars=fars,
                       embed_idxs=np.arange(n_cams))


def run_hugs(steps, batch=4096):
  bind = [f'Config.batch_size = {batch}', 'Config.patch_size = 16', f'Config.image_num_per_batch = {max(1, batch // 256)}', "Config.transient_type = 'withmask'",
          "Config.data_loss_type = 'charb'", 'Config.distortion_loss_mult = 0.001', 'Config.interlevel_loss_mult = 1.0',
          'Model.opaque_background = True', 'Model.num_levels = 2', 'Model.num_prop_samples = 64',
          'Model.num_nerf_samples = 128', 'Model.num_glo_features = 48', 'Model.num_embeddings = 64',
          'PropMLP.net_depth = 4', 'PropMLP.net_width = 256', 'PropMLP.disable_rgb = True', 'NerfMLP.net_depth = 8',
          'NerfMLP.net_width = 256']
  config = configs.load_config([], bind, save_config=False)
  dev = device()
  model, state, _, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=batch, device=dev)
  dd = hugs_dataset()
  gen = torch.Generator(device=dev); gen.manual_seed(1 + RANK)

  def step():
    b = dd.next_train_batch(gen, batch, config.patch_size, config.patch_dilation, config.image_num_per_batch)
    return train_pstep(gen, state, b, min(1.0, (state.step + 1) / config.max_steps), None)

  for _ in range(10):
    step()
  ms, out = timed(step, steps, dev)
  stats = out[1]
  b = dd.next_train_batch(gen, batch, config.patch_size, config.patch_dilation, config.image_num_per_batch)
  emit({'config': 'BASELINE config 3: Mip-NeRF 360 + HuGS static masks (Phototourism-shape synthetic, patch 16, GLO 48, no '
                  'warp / raydist), device-side batch assembly, rays sharded over the GPUs (weak scaling)',
        'metric': 'training rays/s', 'value': WORLD * batch / ms * 1e3, 'ms_per_step': ms, 'rays_per_gpu': batch,
        'n_gpus': WORLD, 'steps': steps, 'loss': float(stats['loss']),
        'static_fraction': float(b.rays.static_mask.mean()), 'h2d_bytes_per_step': 0})


def run_config_b(steps, batch=4096):
  """SURVEY §8d variant B: the repo-default 3-level 64 / 64 / 32 sampling (360.gin geometry, 256-wide MLPs)."""
  import bench
  bind = [b for b in bench.gin_bindings(batch) if 'num_levels' not in b and 'num_nerf_samples' not in b]
  bind += ['Model.num_levels = 3', 'Model.num_nerf_samples = 32']
  config = configs.load_config([], bind, save_config=False)
  dev = device()
  model, state, _, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=batch, device=dev)
  rays, rgb = bench.synthetic_batch(batch, seed=5 + RANK)
  b = utils.Batch(rays=utils.Rays(**{k: v.to(dev) for k, v in rays.items()}), rgb=rgb.to(dev))
  gen = torch.Generator(device=dev); gen.manual_seed(1)
  for _ in range(10):
    train_pstep(gen, state, b, 0.1, None)
  ms, out = timed(lambda: train_pstep(gen, state, b, 0.1, None), steps, dev)
  emit({'config': 'Mip-NeRF 360 variant B: 3 levels, 64 / 64 / 32 samples, 256-wide MLPs (SURVEY §8d)',
        'metric': 'training rays/s', 'value': WORLD * batch / ms * 1e3, 'ms_per_step': ms, 'rays_per_gpu': batch,
        'n_gpus': WORLD, 'steps': steps, 'loss': float(out[1]['loss'])})


def run_render(resolutions=((800, 800), (720, 1280), (1080, 1920), (1440, 2560), (2160, 3840))):
  import bench
  config = configs.load_config([], bench.gin_bindings(4096), save_config=False)
  config.render_chunk_size = 65536 * WORLD          # every rank renders 65536 rays of a chunk
  dev = device()
  model, state, render_eval_pfn, _, _ = train_utils.setup_model(config, rng=0, max_rays=65536, device=dev)
  rng = np.random.default_rng(0)
  for h, w in resolutions:
    p2c, c2w = sphere_cameras(rng, 1, np.array([[h, w]]))
    dd = DeviceDataset(p2c, c2w, [h], [w], near=0.2, far=1e6)
    rays = dd.generate_ray_batch(0).rays
    fn = lambda _, chunk: render_eval_pfn(state.params, 1.0, None, chunk)
    render = lambda: models.render_image(fn, rays, None, config, verbose=False, world_size=WORLD)
    render()                                                                # warm-up
    ms, out = timed(render, 1, dev)
    dt = ms * 1e-3
    assert tuple(out['rgb'].shape) == (h, w, 3) and torch.isfinite(out['rgb']).all()
    emit({'config': f'BASELINE config 5: full-frame render {w}x{h} (config A weights, compute_extras, 65536 rays per GPU '
                    f'and chunk, rows striped over the GPUs, all-gather per chunk)',
          'metric': 'render rays/s', 'value': h * w / dt, 'frame_s': dt, 'n_gpus': WORLD,
          'outputs': sorted(k for k in out if not k.startswith('ray_'))})


def run_render_frame(resolutions=((800, 800), (720, 1280), (1080, 1920), (1440, 2560), (2160, 3840))):
  """Config 5 through the frame pipeline (models.render_frame -> hugs_render_frame): ray generation on the device INSIDE the timed
  region, one library call per frame and rank, uint8 quantisation + squared error on the device, one all-gather per frame."""
  import bench
  config = configs.load_config([], bench.gin_bindings(4096), save_config=False)
  dev = device()
  model, state, _, _, _ = train_utils.setup_model(config, rng=0, max_rays=65536, device=dev)
  rng = np.random.default_rng(0)
  for h, w in resolutions:
    p2c, c2w = sphere_cameras(rng, 1, np.array([[h, w]]))
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    dd = DeviceDataset(p2c, c2w, [h], [w], images=[img], near=0.2, far=1e6)
    render = lambda: models.render_frame(model, state.params, dd, 0, 1.0, config, compute_extras=True, want_u8=True,
                                         want_psnr=True)
    render()                                                                # warm-up
    ms, out = timed(render, 2, dev)
    dt = ms * 1e-3
    assert tuple(out['rgb'].shape) == (h, w, 3) and torch.isfinite(out['rgb']).all() and np.isfinite(out['psnr'])
    emit({'config': f'BASELINE config 5: full-frame render {w}x{h} through hugs_render_frame (config A weights, compute_extras, '
                    f'on-device ray generation, uint8 quantisation and PSNR; rows striped over the GPUs, one all-gather per frame)',
          'metric': 'render rays/s', 'value': h * w / dt, 'frame_s': dt, 'n_gpus': WORLD, 'psnr': out['psnr'],
          'outputs': sorted(out.keys())})


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('what', nargs='*', default=['hugs', 'render'])
  ap.add_argument('--steps', type=int, default=100)
  a = ap.parse_args()
  if 'hugs' in a.what:
    run_hugs(a.steps)
  if 'render' in a.what:
    run_render()
  if 'frame' in a.what:
    run_render_frame()
  if 'b' in a.what:
    run_config_b(a.steps)
  if WORLD > 1:
    dist.barrier()
    dist.destroy_process_group()
