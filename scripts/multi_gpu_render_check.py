"""torchrun --nproc-per-node N scripts/multi_gpu_render_check.py: models.render_image with rays striped over N ranks
(create_render_fn all-gathers the shards) must equal a single-rank render of the same frame; also runs two training
steps and checks that the replicas' parameters stay bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from nerf_hugs_b200.internal import configs, models, train_utils, utils
from tests import helpers as H

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
bind = ['Config.batch_size = 512', 'Config.near = 0.2', 'Config.far = 1e6', 'Config.render_chunk_size = 256',
        'Model.raydist_fn = @jnp.reciprocal', 'Model.opaque_background = True', 'Model.num_levels = 2',
        'Model.num_prop_samples = 64', 'Model.num_nerf_samples = 128', 'PropMLP.warp_fn = @coord.contract',
        'PropMLP.net_depth = 4', 'PropMLP.disable_rgb = True', 'NerfMLP.warp_fn = @coord.contract', 'NerfMLP.net_width = 256']
config = configs.load_config([], bind, save_config=False)
model, state, render_eval_pfn, train_pstep, lr_fn = train_utils.setup_model(config, rng=0, max_rays=1024, device=dev)

# training: every rank takes its slice of one global batch; parameters must stay replicated
rays, gt = H.make_rays(512, seed=1)
mine = utils.Batch(rays=utils.Rays(**{k: utils.rank_slice(v, rank, world) for k, v in rays.items()}),
                   rgb=utils.rank_slice(gt, rank, world))
gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
# the update train_pstep must produce in its first step: pmean of the per-rank gradients (train_utils.py:457-458), then
# clip + Adam - computed here with ONE plain all-reduce, against train_pstep's two overlapped collectives
from nerf_hugs_b200 import _lib
eng = model.engine
p0 = state.params.clone()
eng.params_changed(p0)
eng.set_train_rng(train_utils._rng_seed(gen), 0)
g, _ = eng.loss_and_grad(p0, {k: v.to(dev) for k, v in mine.rays.as_dict().items()}, mine.rgb.to(dev), 0.1, None,
                         train_utils.loss_cfg_from(config))
g = g.clone()
dist.all_reduce(g)
a = _lib.AdamCfg()
a.lr, a.beta1, a.beta2, a.eps = float(lr_fn(0)), config.adam_beta1, config.adam_beta2, config.adam_eps
a.grad_max_norm, a.grad_max_val, a.step, a.grad_scale = config.grad_max_norm, config.grad_max_val, 0, 1.0 / world
p_expect, mu, nu = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
eng.adam_step(p_expect, g, mu, nu, a)
eng.params_changed(state.params)
for i in range(2):
  state, stats, gen = train_pstep(gen, state, mine, 0.1, None)
  if i == 0:
    d_exp, d_got = (p_expect - p0), (state.params - p0)
    rel = float((d_exp - d_got).norm() / d_exp.norm())
    assert rel < 1e-3, f'overlapped all-reduce changed the update: {rel}'      # fp32 atomics reorder the reductions
ref = state.params.clone()
dist.broadcast(ref, 0)
assert torch.equal(ref, state.params), 'replicas diverged'

# rendering: H x W frame striped over ranks vs the same frame rendered by this rank alone
Hh, Ww = 13, 37
r2, _ = H.make_rays(Hh * Ww, seed=3)
img = utils.Rays(**{k: v.reshape(Hh, Ww, -1) for k, v in r2.items()})
out = models.render_image(lambda rng, rr: render_eval_pfn(state.params, 0.5, None, rr), img, None, config,
                          verbose=False, world_size=world)
solo, _ = model.apply(state.params, None, r2, 0.5, True)
err = float((out['rgb'].reshape(-1, 3) - solo[-1]['rgb']).abs().max())
assert out['rgb'].shape == (Hh, Ww, 3) and err < 1e-6, err
dist.barrier()
if rank == 0:
  print(f'multi-GPU check ok: world {world}, loss {stats["loss"]:.5f}, striped render max |diff| {err:.2e}')
dist.destroy_process_group()
