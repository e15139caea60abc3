#!/usr/bin/env python
"""Benchmark of the north-star metric: training rays/s of the Mip-NeRF 360 per-ray path
(4096-ray batch, 64 proposal + 128 NeRF samples per ray, 256-wide MLPs; SURVEY.md §8d config 2 / "A").

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, tcgen05 path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

One step = one optimisation step (sampling -> IPE -> PropMLP/NerfMLP -> compositing -> losses ->
backward -> gradient all-reduce (N > 1) -> clip + Adam) on one synthetic batch: rays of 100 cameras on a
sphere (800x800, focal 1111), uniform-random target colours, random-init weights (no dataset / checkpoint
is available offline).  Rays are sharded over ranks (one process per GPU, weak scaling: 4096 rays per GPU,
`--scaling strong` keeps the global batch at 4096).

Printed JSON (rank 0, one line): see the keys below; `value` = whole-job rays/s with inputs resident in HBM,
`e2e` = the same metric through the reference-facing call (`train_utils.train_pstep`) with HOST batches
(pinned H2D copy of every batch + D2H read of the loss inside the timed region),
`roofline` = the kernel class with the largest time per step against its measured peak (tensor or HBM), with every
MLP kernel listed under `roofline.kernels`,
`cpu_baseline` = the CPU oracle (a port of the reference algorithm) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVELS, N_PROP, N_NERF, WIDTH, NERF_WIDTH = 2, 64, 128, 256, 256
FLOPS_PROP_SAMPLE = 651776          # 2*sum(K*N) over the PropMLP(256) Dense layers


def nerf_sample_flops(w):
  """2 * sum(K * N) over the NerfMLP Dense layers: 1,638,400 at width 256, 17,344,000 at 1024."""
  return 2 * (504 * w + 3 * w * w + w * w + (w + 504) * w + 2 * w * w + w + w * 256 + 283 * 128 + 128 * 3)


FLOPS_NERF_SAMPLE = nerf_sample_flops(NERF_WIDTH)
FLOPS_RAY_FWD = (LEVELS - 1) * N_PROP * FLOPS_PROP_SAMPLE + N_NERF * FLOPS_NERF_SAMPLE   # config A: 251.4 MFLOP
METRIC = 'training rays/s (Mip-NeRF 360, 4096-ray batch, 64+128 samples/ray, 256-wide MLPs)'
WORKLOAD = ('Mip-NeRF 360 config A (SURVEY §8d): 360.gin geometry, num_levels=2, 64 proposal + 128 NeRF samples/ray, '
            'PropMLP 4x256, NerfMLP 8x256, IPE 504, contract + reciprocal spacing')
WORKLOADS = {   # BASELINE config 2 and its variants: levels, proposal samples, NeRF samples, NerfMLP width
    'A': (2, 64, 128, 256, None),
    'Aprime': (2, 64, 128, 1024, "variant A' of config A: NerfMLP.net_width = 1024 (every shipped gin), layer-at-a-time path"),
    'B': (3, 64, 32, 256, 'variant B: the repo-default 3 levels 64 / 64 / 32, NerfMLP 8x256'),
    '360gin': (3, 64, 32, 1024, 'MipNeRF360/configs/360.gin as shipped: 3 levels 64 / 64 / 32, NerfMLP 8x1024'),
}


def set_workload(name):
  global LEVELS, N_PROP, N_NERF, NERF_WIDTH, FLOPS_NERF_SAMPLE, FLOPS_RAY_FWD, WORKLOAD
  LEVELS, N_PROP, N_NERF, NERF_WIDTH, note = WORKLOADS[name]
  FLOPS_NERF_SAMPLE = nerf_sample_flops(NERF_WIDTH)
  FLOPS_RAY_FWD = (LEVELS - 1) * N_PROP * FLOPS_PROP_SAMPLE + N_NERF * FLOPS_NERF_SAMPLE
  if note:
    WORKLOAD = f'Mip-NeRF 360, {note}; 360.gin geometry (contract + reciprocal spacing), IPE 504, PropMLP 4x256'

# dram__bytes_read.sum + dram__bytes_write.sum per launch come from the newest committed `ncu --set full` summary under
# profiles/ (written by scripts/ncu_summary.py from a capture of THIS workload), never from constants in this file
NCU_DRAM_FILE = os.path.join(ROOT, 'profiles', 'ncu_dram_bytes.json')


def ncu_dram_bytes():
  try:
    d = json.load(open(NCU_DRAM_FILE))
    return d.get('bytes_per_launch', {}), d.get('source')
  except Exception:
    return {}, None


def workload_config(args, per_gpu, global_batch):
  """`config` of the JSON line: identical for the CUDA arm and the reference arm (the workload, not the implementation)."""
  return {'workload': WORKLOAD, 'rays_per_gpu': per_gpu, 'global_batch': global_batch, 'scaling': args.scaling}


def synthetic_batch(n_rays, seed, n_cams=100, hw=800, focal=1111.1):
  """Uniform random (camera, pixel) draws from cameras on the unit sphere looking at the origin."""
  import torch
  rng = np.random.default_rng(seed)
  cam_pos = rng.normal(size=(n_cams, 3)); cam_pos /= np.linalg.norm(cam_pos, axis=-1, keepdims=True)
  fwd = -cam_pos
  up = np.array([0., 0., 1.])
  right = np.cross(fwd, up); right /= np.linalg.norm(right, axis=-1, keepdims=True) + 1e-9
  upv = np.cross(right, fwd)
  ci = rng.integers(0, n_cams, size=n_rays)
  px = rng.uniform(0, hw, size=(n_rays, 2))
  x = (px[:, 0] - hw / 2) / focal; y = -(px[:, 1] - hw / 2) / focal
  d = fwd[ci] + x[:, None] * right[ci] + y[:, None] * upv[ci]
  v = d / np.linalg.norm(d, axis=-1, keepdims=True)
  f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
  rays = dict(origins=f32(cam_pos[ci]), directions=f32(d), viewdirs=f32(v),
              radii=torch.full((n_rays, 1), float(1.0 / focal * 2 / np.sqrt(12))),
              near=torch.full((n_rays, 1), 0.2), far=torch.full((n_rays, 1), 1e6),
              lossmult=torch.ones(n_rays, 1), static_mask=torch.ones(n_rays, 1),
              embed_idx=torch.tensor(ci[:, None], dtype=torch.int32))
  rgb = f32(rng.uniform(size=(n_rays, 3)))
  return rays, rgb


def gin_bindings(batch_size):
  """Explicit bindings of benchmark config A."""
  return [f'Config.batch_size = {batch_size}', 'Config.near = 0.2', 'Config.far = 1e6', 'Config.patch_size = 1',
          "Config.data_loss_type = 'charb'", 'Config.distortion_loss_mult = 0.01', 'Config.interlevel_loss_mult = 1.0',
          'Model.raydist_fn = @jnp.reciprocal', 'Model.opaque_background = True', f'Model.num_levels = {LEVELS}',
          f'Model.num_prop_samples = {N_PROP}', f'Model.num_nerf_samples = {N_NERF}',
          'PropMLP.warp_fn = @coord.contract', 'PropMLP.net_depth = 4', f'PropMLP.net_width = {WIDTH}',
          'PropMLP.disable_rgb = True', 'NerfMLP.warp_fn = @coord.contract', 'NerfMLP.net_depth = 8',
          f'NerfMLP.net_width = {NERF_WIDTH}']


class ClockSampler(threading.Thread):
  """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.rows, self._halt = index, [], threading.Event()

  def run(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not self._halt.is_set():
      try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
          self.rows.append([c.strip() for c in out.split(',')])
      except Exception:
        pass
      self._halt.wait(0.05)

  def stop(self):
    self._halt.set()
    self.join(timeout=5)
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower() == 'active'})
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': reasons, 'samples': len(sm)}


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return {'bf16_sustained': p.get('bf16_tflops_sustained'), 'bf16_burst': p.get('bf16_tflops'),
            'hbm_gbs': p.get('hbm_gbs'), 'source': 'measured'}
  return {'bf16_sustained': 1400.0, 'bf16_burst': 1590.0, 'hbm_gbs': 6650.0, 'source': 'fallback'}


def cpu_oracle_rate(sample_rays, repeats, threads=None):
  """The CPU port of the reference algorithm (oracle/) on a bounded sample: train-step rays/s."""
  import torch
  from oracle import mipnerf360 as O
  if threads:
    torch.set_num_threads(threads)
  ocfg = O.ModelConfig(num_levels=LEVELS, num_prop_samples=N_PROP, num_nerf_samples=N_NERF, raydist_fn='reciprocal',
                       opaque_background=True,
                       nerf_mlp=O.MLPConfig(net_depth=8, net_width=NERF_WIDTH, warp_fn='contract'),
                       prop_mlp=O.MLPConfig(net_depth=4, net_width=WIDTH, disable_rgb=True, warp_fn='contract'))
  lcfg = O.LossConfig()
  basis = torch.tensor(np.load(os.path.join(ROOT, 'tests', 'golden', 'geopoly_basis.npz'))['icosahedron_2'].T,
                       dtype=torch.float32)
  params = O.init_params(ocfg, seed=0)
  opt = O.init_opt_state(params)
  rays, rgb = synthetic_batch(sample_rays, seed=123)
  jit = [torch.rand(sample_rays, 1) for _ in range(LEVELS)]
  times = []
  for i in range(repeats + 1):
    t0 = time.perf_counter()
    params, opt, _, _ = O.train_step(ocfg, lcfg, params, opt, i, rays, rgb, 0.5, basis, jitter=jit)
    times.append(time.perf_counter() - t0)
  dt = float(np.mean(times[1:]))
  return sample_rays / dt, dt, torch.get_num_threads()


def run_reference(args):
  """--impl reference: the reference algorithm on the host cores (the JAX stack is not installable offline,
  so this is the CPU port in oracle/ — the only other place bench.py executes oracle code)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import torch
  sample = args.ref_rays
  from oracle import mipnerf360 as O  # noqa: F401  (import cost outside the timed region)
  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  world = int(os.environ.get('WORLD_SIZE', '1'))
  per_gpu = args.rays if args.scaling == 'weak' else args.rays // world
  # one warm-up pass inside cpu_oracle_rate, then `steps` timed repeats (warm-ups beyond 1 add nothing on CPU)
  rate, dt, threads = cpu_oracle_rate(sample, max(1, args.steps), threads=cores)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'rays/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
      'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(args, per_gpu, per_gpu * world),
      'cpu_baseline': {'value': rate, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
                       'sample': f'each step = a bounded sample of {sample} rays of the batch, {args.steps} train '
                                 f'steps (fwd+bwd+Adam) of the torch fp32 port of the reference JAX path (JAX is not '
                                 f'installable offline), torch threads = os.cpu_count() = {cores}'},
      'e2e': {'value': rate, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))


def run_ours(args):
  import torch
  import torch.distributed as dist
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (B200); the product path has no CPU fallback')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    # NCCL announces its version on stdout when the first communicator is created: keep stdout for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
      dist.init_process_group('nccl', device_id=dev)
      dist.all_reduce(torch.zeros(1, device=dev))
      torch.cuda.synchronize()
    finally:
      sys.stdout.flush()
      os.dup2(saved_stdout, 1)
      os.close(saved_stdout)
  from nerf_hugs_b200.internal import configs, train_utils, utils
  from nerf_hugs_b200.engine import Engine

  per_gpu = args.rays if args.scaling == 'weak' else args.rays // world
  global_batch = per_gpu * world
  config = configs.load_config([], gin_bindings(global_batch), save_config=False)
  model, state, render_eval_pfn, train_pstep, lr_fn = train_utils.setup_model(config, rng=0, max_rays=per_gpu, device=dev)
  eng = model.engine
  gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)

  # a pool of distinct host batches (pinned) and their device copies
  n_pool = 4
  host, devb = [], []
  for i in range(n_pool):
    rays, rgb = synthetic_batch(per_gpu, seed=1000 * rank + i)
    rays_p = {k: v.pin_memory() for k, v in rays.items()}
    rgb_p = rgb.pin_memory()
    host.append(utils.Batch(rays=utils.Rays(**rays_p), rgb=rgb_p))
    devb.append(utils.Batch(rays=utils.Rays(**{k: v.to(dev) for k, v in rays.items()}), rgb=rgb.to(dev)))
  h2d = sum(v.numel() * v.element_size() for v in host[0].rays.as_dict().values()) + host[0].rgb.numel() * 4

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(batches, steps, read_loss):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss, prev = 0.0, None
    for i in range(steps):
      _, stats, _ = train_pstep(gen, state, batches[i % len(batches)], min(1.0, (state.step + 1) / config.max_steps), None)
      if read_loss:
        # device -> host read of every step's result: train_pstep copies the stats to pinned memory asynchronously,
        # the host consumes step i-1's loss while step i runs (and the last one before the timer stops)
        if prev is not None:
          loss = prev['loss']
        prev = stats
    if read_loss and prev is not None:
      loss = prev['loss']
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), loss

  W = max(args.warmup, 3)
  timed(devb, W, False)
  launches0 = Engine.launch_count()
  sampler = ClockSampler(local) if rank == 0 else None
  if sampler:
    sampler.start()
  ms, _ = timed(devb, args.steps, False)
  clocks = sampler.stop() if sampler else None
  launches = Engine.launch_count() - launches0
  value = global_batch * args.steps / (ms * 1e-3)

  # end-to-end: host batches through the reference-facing call, loss read back every step
  timed(host, 3, True)
  ms_e2e, last_loss = timed(host, args.steps, True)
  e2e = global_batch * args.steps / (ms_e2e * 1e-3)

  # per-kernel-class CUDA-event timing (separate short pass so the events do not perturb `value`)
  eng.profile(True)
  prof_steps = min(args.steps, 50)
  timed(devb, prof_steps, False)
  prof = eng.profile_read()
  eng.profile(False)
  kern_ms = {k: v[0] / prof_steps for k, v in prof.items() if v[1] > 0}

  if rank == 0:
    peaks = measured_peaks()
    src = f"{peaks['source']} (MEASURED_PEAKS.json)"
    n_nerf, n_prop = per_gpu * N_NERF, per_gpu * N_PROP * (LEVELS - 1)
    wn = NERF_WIDTH

    def tensor_line(cls, kernel, flops):
      t = kern_ms.get(cls)
      a = flops / (t * 1e-3) / 1e12 if t else None
      return {'class': cls, 'kernel': kernel, 'bound': 'tensor', 'achieved': a, 'peak': peaks['bf16_sustained'],
              'unit': 'TFLOP/s', 'frac': a / peaks['bf16_sustained'] if a else None,
              'algorithmic_flops_per_launch': flops, 'ms_per_launch': t,
              'peak_source': 'cuBLAS bf16 sustained, ' + src}

    def hbm_line(cls, kernel, nbytes):
      t = kern_ms.get(cls)
      a = nbytes / (t * 1e-3) / 1e9 if t else None
      return {'class': cls, 'kernel': kernel, 'bound': 'hbm', 'achieved': a, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
              'frac': a / peaks['hbm_gbs'] if a else None, 'algorithmic_bytes_per_launch': nbytes, 'ms_per_launch': t,
              'peak_source': 'STREAM-style copy, ' + src}

    # algorithmic work per launch (DESIGN.md "Kernels and their bounds"): MLP chains = SURVEY §8d FLOPs per sample;
    # the weight-gradient GEMMs stream every saved activation / dZ / feature row exactly once (unique bytes).
    wg_bytes_nerf = n_nerf * ((8 + 2) * 512 * 2 + 1024 + 128)
    wg_bytes_prop = n_prop * (4 * 512 * 2 + 1024 + 128)
    wide = wn != 256    # layer-at-a-time path (dense_tc.cu / layered.cu) instead of the chain kernel
    lines = [
        tensor_line('chain_fwd_nerf', 'dense_tc_kernel NerfMLP forward GEMMs, one launch per layer (tcgen05 cta_group::2)' if wide
                    else 'mlp_pp_kernel<train, cta_pair> NerfMLP forward chain (tcgen05 cta_group::2)',
                    n_nerf * FLOPS_NERF_SAMPLE),
        tensor_line('chain_bwd_nerf', 'dense_tc_kernel NerfMLP dgrad GEMMs (+ bias column sums)' if wide
                    else 'mlp_pp_kernel<train, cta_pair> NerfMLP dgrad chain',
                    n_nerf * 2 * (7 * wn * wn + wn * 256 + 256 * 128)),
        (hbm_line('wgrad_nerf', 'wgrad_kernel NerfMLP weight gradients (tcgen05, MN-major operands)', wg_bytes_nerf)
         if not wide else tensor_line('wgrad_nerf', 'wgrad2_kernel NerfMLP weight gradients on CTA pairs, one launch per layer '
                                      '(wgrad_kernel with HUGS_WGRAD_PAIRS=0)', n_nerf * FLOPS_NERF_SAMPLE)),
        tensor_line('chain_fwd_prop', 'mlp_pp_kernel<train, cta_pair> PropMLP forward chain', n_prop * FLOPS_PROP_SAMPLE),
        hbm_line('wgrad_prop', 'wgrad_kernel PropMLP weight gradients', wg_bytes_prop),
    ]
    # the per-ray kernels (HBM / latency bound): algorithmic bytes = what one launch must read and write
    samp_bytes = per_gpu * ((2 * (N_PROP + 1) * 4 + 8) + ((2 * N_PROP + 1) * 4 + 2 * (N_NERF + 1) * 4 + 8))
    comp_bytes = per_gpu * (N_PROP * 8 + (N_PROP + 1) * 4 + 12                                   # proposal composite
                            + N_NERF * 36 + 2 * (N_NERF + 1) * 4 + 28                            # final loss + backward
                            + N_PROP * 8 + 2 * (N_PROP + 1) * 4 + N_NERF * 4 + (N_NERF + 1) * 4)  # interlevel + backward
    if LEVELS == 2:
      lines += [hbm_line('sample', 'resample_kernel x2 (dilate + resample, one warp per ray)', samp_bytes),
                hbm_line('composite_loss', 'composite_kernel + final_loss_bwd_kernel + prop_loss_bwd_kernel + reductions',
                         comp_bytes)]
    lines = [l for l in lines if l['ms_per_launch']]
    # the weight gradients of the 256-wide path are HBM-bound by construction; the same launch on the tensor lens
    # (SURVEY §8d FLOPs of the layers it differentiates) is reported next to it
    for l in lines:
      if l['class'] in ('wgrad_nerf', 'wgrad_prop') and l['bound'] == 'hbm':
        fl = (n_nerf * FLOPS_NERF_SAMPLE) if l['class'] == 'wgrad_nerf' else (n_prop * FLOPS_PROP_SAMPLE)
        l['tensor_lens'] = {'achieved': fl / (l['ms_per_launch'] * 1e-3) / 1e12, 'unit': 'TFLOP/s',
                            'frac': fl / (l['ms_per_launch'] * 1e-3) / 1e12 / peaks['bf16_sustained']}
    dom = max(lines, key=lambda l: l['ms_per_launch'])
    dram, dram_src = ncu_dram_bytes()
    for l in lines:
      # the ncu capture is of config A (chain kernels, 4096 rays): other configs carry no measured traffic
      l['traffic'] = dram.get(l['class']) if (per_gpu == 4096 and not wide and LEVELS == 2) else None
    roofline = dict(dom)
    roofline['traffic_source'] = dram_src
    roofline['dominant_by'] = 'largest CUDA-event time per step among the kernel classes'
    roofline['step_tflops_3x_convention'] = 3 * FLOPS_RAY_FWD * per_gpu / (ms / args.steps * 1e-3) / 1e12
    roofline['step_frac_of_tensor_peak'] = roofline['step_tflops_3x_convention'] / peaks['bf16_sustained']
    roofline['kernels'] = lines
    roofline['kernel_class_ms_per_step'] = kern_ms
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
      rate, dt, threads = cpu_oracle_rate(args.ref_rays, 3)
      cpu = {'value': rate, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
             'sample': f'{args.ref_rays} rays x 3 train steps (fwd+bwd+Adam) of the same workload, torch fp32 oracle '
                       f'(port of the reference JAX path), os.cpu_count()={os.cpu_count()}'}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': workload_config(args, per_gpu, global_batch),
        'arm': {'parallelism': f'ray-sharded dp{world}, one process per GPU, NerfMLP gradient all-reduce overlapped with '
                               f'the proposal backward',
                'l2_policy': 'per-step working set (>5 GB of saved activations) >> 126 MB L2; 4 distinct batches cycled',
                'precision': 'bf16 operands, fp32 accumulate (tcgen05); fp32 sampling/compositing/losses/Adam'},
        'e2e': {'value': e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 64 + 36,
                'ms_per_step': ms_e2e / args.steps, 'last_loss': float(last_loss)},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
    }
    if cpu:
      line['cpu_baseline'] = cpu
    print(json.dumps(line))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--rays', type=int, default=4096, help='rays per GPU (weak) or global batch (strong)')
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
  ap.add_argument('--ref-rays', type=int, default=128, help='bounded CPU sample (rays per step)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--config', default='A', choices=sorted(WORKLOADS), help='A = BASELINE config 2 (default); variants of SURVEY §8d')
  args = ap.parse_args()
  set_workload(args.config)
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
