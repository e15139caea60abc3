"""The reference-facing call surface on the GPU: configs -> train_utils.setup_model -> train_pstep ->
models.render_image, as MipNeRF360/train.py and eval.py drive it (train.py:51-142, eval.py:95-104)."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _config(batch, glo=0, transient=None):
  from nerf_hugs_b200.internal import configs
  b = [f'Config.batch_size = {batch}', 'Config.near = 0.2', 'Config.far = 1e6', 'Config.render_chunk_size = 256',
       'Model.raydist_fn = @jnp.reciprocal', 'Model.opaque_background = True', 'Model.num_levels = 2',
       'Model.num_prop_samples = 64', 'Model.num_nerf_samples = 128', f'Model.num_glo_features = {glo}',
       'Model.num_embeddings = 16', 'PropMLP.warp_fn = @coord.contract', 'PropMLP.net_depth = 4',
       'PropMLP.disable_rgb = True', 'NerfMLP.warp_fn = @coord.contract', 'NerfMLP.net_width = 256']
  if transient:
    b.append(f"Config.transient_type = '{transient}'")
  return configs.load_config([], b, save_config=False)


@pytest.mark.parametrize('glo,transient', [(0, None), (4, 'withmask')])
def test_setup_model_train_and_render(glo, transient):
  from nerf_hugs_b200.internal import models, train_utils, utils
  config = _config(128, glo, transient)
  model, state, render_eval_pfn, train_pstep, lr_fn = train_utils.setup_model(config, rng=0, max_rays=512)
  assert abs(lr_fn(0) - 2e-5) < 1e-12 and state.params.numel() == model.engine.n_params
  tree = state.tree(model)['params']
  assert tree['NerfMLP_0']['Dense_0']['kernel'].shape == (504, 256)
  assert tree['PropMLP_0']['Dense_4']['kernel'].shape == (256, 1)
  rays, gt = H.make_rays(128, seed=2)
  batch = utils.Batch(rays=utils.Rays(**{k: v.pin_memory() for k, v in rays.items()}), rgb=gt.pin_memory())
  gen = torch.Generator(device=model.engine.device); gen.manual_seed(0)
  losses = []
  for step in range(8):
    state, stats, gen = train_pstep(gen, state, batch, (step + 1) / 100, None)
    losses.append(stats['loss'])
  assert state.step == 8 and all(np.isfinite(losses)) and losses[-1] < losses[0]
  for k in ('loss', 'losses', 'mses', 'psnrs', 'psnr', 'grad_norms', 'grad_maxes'):
    assert k in stats.keys()
  assert set(stats['losses']) == {'data', 'interlevel', 'distortion'}
  # full-frame render through render_image with the reference's chunk / shard / gather plumbing
  H_, W_ = 9, 31
  r2, _ = H.make_rays(H_ * W_, seed=3)
  img_rays = utils.Rays(**{k: v.reshape(H_, W_, -1) for k, v in r2.items()})
  rendering = models.render_image(lambda rng, rr: render_eval_pfn(state.params, 0.5, None, rr), img_rays, None,
                                  config, verbose=False)
  assert rendering['rgb'].shape == (H_, W_, 3) and rendering['distance_median'].shape == (H_, W_)
  assert len(rendering['ray_sdist']) == 2 and rendering['ray_sdist'][1].shape == (config.vis_num_rays, 129)
  direct, _ = model.apply(state.params, None, r2, 0.5, True, zero_glo=False)
  np.testing.assert_allclose(rendering['rgb'].reshape(-1, 3).cpu().numpy(), direct[-1]['rgb'].cpu().numpy(),
                             rtol=0, atol=1e-6)


def test_unsupported_transient_type_is_loud():
  from nerf_hugs_b200.internal import train_utils
  config = _config(128, transient='robustnerf')
  with pytest.raises(NotImplementedError, match='robustnerf'):
    train_utils.setup_model(config, rng=0)


def test_checkpoint_round_trip_through_the_flax_format(tmp_path):
  """train.py:121,235: save_checkpoint(state) -> restore_checkpoint into a fresh setup_model gives the same renders."""
  from nerf_hugs_b200.internal import checkpoints, train_utils, utils
  config = _config(64, glo=4)
  model, state, render_eval_pfn, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=128)
  rays, gt = H.make_rays(64, seed=2, glo=True)
  batch = utils.Batch(rays=utils.Rays(**rays), rgb=gt)
  gen = torch.Generator(device=model.engine.device); gen.manual_seed(0)
  for _ in range(2):
    state, _, gen = train_pstep(gen, state, batch, 0.1, None)
  checkpoints.save_checkpoint(str(tmp_path), state, state.step, model=model, keep=100)
  sd = checkpoints.restore_checkpoint(str(tmp_path), None)
  assert int(sd['step']) == 2 and 'kernel' in sd['params']['params']['NerfMLP_0']['Dense_0']
  assert sd['params']['params']['GloEmbed_0']['embedding'].shape == (16, 4)
  model2, state2, _, _, _ = train_utils.setup_model(config, rng=1, max_rays=128)
  assert not torch.equal(state2.params, state.params)
  state2 = checkpoints.restore_checkpoint(str(tmp_path), state2, model=model2)
  assert state2.step == 2 and torch.equal(state2.params, state.params) and torch.equal(state2.mu, state.mu)
  a, _ = model.apply(state.params, None, utils.Rays(**rays), 0.5, True)
  b, _ = model2.apply(state2.params, None, utils.Rays(**rays), 0.5, True)
  assert torch.equal(a[-1]['rgb'], b[-1]['rgb'])


def test_render_only_handle_does_not_reserve_training_buffers():
  """Saved activations / dZ / gates (10.6 KB per NeRF sample) are allocated by the first training call only: a render
  handle sized for a 16384-ray chunk stays small; the same handle can still start training later."""
  from nerf_hugs_b200.internal import train_utils, utils
  torch.cuda.synchronize(); torch.cuda.empty_cache()
  free0, _ = torch.cuda.mem_get_info()
  config = _config(16384)
  model, state, render_eval_pfn, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=16384)
  rays, gt = H.make_rays(256, seed=4)
  model.apply(state.params, None, utils.Rays(**rays), 0.5, True)
  torch.cuda.synchronize()
  free1, _ = torch.cuda.mem_get_info()
  used_render = free0 - free1
  assert used_render < 8 * 2 ** 30, used_render          # features (3 GB) + per-sample fp32 buffers, not the 27 GB of saves
  gen = torch.Generator(device=model.engine.device); gen.manual_seed(0)
  state, stats, gen = train_pstep(gen, state, utils.Batch(rays=utils.Rays(**rays), rgb=gt), 0.1, None)
  assert np.isfinite(stats['loss'])
  free2, _ = torch.cuda.mem_get_info()
  assert free1 - free2 > 12 * 2 ** 30                    # the training buffers appeared with the first training step
  model.engine.close()


def test_train_pstep_stats_contract():
  """stats of train_step (train_utils.py:442-476): every key train.py:174-213 histograms, with summarize_tree's three
  key depths, checked against values recomputed from the parameters / gradient / update themselves."""
  import math
  from nerf_hugs_b200 import _lib
  from nerf_hugs_b200.internal import train_utils, utils
  config = _config(64, glo=4)
  model, state, _, train_pstep, lr_fn = train_utils.setup_model(config, rng=0, max_rays=128)
  eng = model.engine
  rays, gt = H.make_rays(64, seed=2, glo=True)
  batch = utils.Batch(rays=utils.Rays(**rays), rgb=gt)
  p0 = state.params.clone()
  # the gradient of the very same step (deterministic sampling so that it can be recomputed)
  config.randomized = False
  train_pstep = train_utils.create_train_step(model, config)
  eng.params_changed(p0)
  eng.set_train_rng(0, 0)
  g_ref, _ = eng.loss_and_grad(p0, rays, gt, 0.1, None, train_utils.loss_cfg_from(config))
  g_ref = g_ref.clone()
  state, stats, _ = train_pstep(None, state, batch, 0.1, None)
  for k in ('loss', 'losses', 'mses', 'psnrs', 'psnr', 'weight_l2s', 'grad_norms', 'grad_maxes', 'opt_update_norms',
            'opt_update_maxes'):
    assert k in stats, k
  assert len(stats['mses']) == 2 and len(stats['psnrs']) == 2 and np.all(np.isfinite(stats['psnrs']))
  assert 0.0 < stats['mses'][0] < 2.0                     # proposal level: MSE of its background-only rendering
  delta = (state.params - p0)
  for name, off, r, c, _ in eng.layout:
    sl = slice(off, off + r * c)
    np.testing.assert_allclose(stats['weight_l2s'][name], float((p0[sl].double() ** 2).sum()), rtol=1e-5)
    np.testing.assert_allclose(stats['grad_norms'][name], float(g_ref[sl].double().norm()), rtol=1e-3, atol=1e-12)
    np.testing.assert_allclose(stats['grad_maxes'][name], float(g_ref[sl].abs().max()), rtol=1e-3, atol=1e-12)
    np.testing.assert_allclose(stats['opt_update_norms'][name], float(delta[sl].double().norm()), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(stats['opt_update_maxes'][name], float(delta[sl].abs().max()), rtol=1e-6, atol=1e-12)
  # summarize_tree depths: module and layer entries aggregate their tensors
  mods = {'NerfMLP_0', 'PropMLP_0', 'GloEmbed_0'}
  assert mods <= set(stats['weight_l2s']) and 'NerfMLP_0/Dense_3' in stats['grad_norms']
  nerf = [(o, r * c) for n, o, r, c, _ in eng.layout if n.startswith('NerfMLP_0/')]
  tot = sum(float((p0[o:o + k].double() ** 2).sum()) for o, k in nerf)
  np.testing.assert_allclose(stats['weight_l2s']['NerfMLP_0'], tot, rtol=1e-5)
  np.testing.assert_allclose(stats['grad_norms']['NerfMLP_0'],
                             math.sqrt(sum(float((g_ref[o:o + k].double() ** 2).sum()) for o, k in nerf)), rtol=1e-3)
  eng.close()


def test_weight_decay_mults_is_loud():
  from nerf_hugs_b200.internal import train_utils
  config = _config(64)
  config.weight_decay_mults = {'NerfMLP_0': 0.1}
  with pytest.raises(NotImplementedError, match='weight_decay_mults'):
    train_utils.setup_model(config, rng=0, max_rays=128)


def test_apply_with_two_parameter_trees_back_to_back():
  """Model.apply(tree) builds a temporary flat tensor; the packed bf16 operands must follow the tree that
  is being rendered, also when the allocator hands the next temporary the same address."""
  from nerf_hugs_b200.internal import train_utils, utils
  config = _config(64)
  model, state, _, _, _ = train_utils.setup_model(config, rng=0, max_rays=128)
  rays, _ = H.make_rays(64, seed=2)
  r = utils.Rays(**rays)
  t1, t2 = model.init(1), model.init(2)
  a1, _ = model.apply(t1, None, r, 0.5, False)
  a2, _ = model.apply(t2, None, r, 0.5, False)
  b2, _ = model.apply(model.flat_params(t2), None, r, 0.5, False)
  b1, _ = model.apply(model.flat_params(t1), None, r, 0.5, False)
  assert torch.equal(a1[-1]['rgb'], b1[-1]['rgb']) and torch.equal(a2[-1]['rgb'], b2[-1]['rgb'])
  assert not torch.equal(a1[-1]['rgb'], a2[-1]['rgb'])
  model.engine.close()


def test_embed_idx_out_of_range_is_loud_and_memory_safe():
  from nerf_hugs_b200.internal import train_utils, utils
  config = _config(64, glo=4)
  model, state, _, train_pstep, _ = train_utils.setup_model(config, rng=0, max_rays=128)
  rays, gt = H.make_rays(64, seed=2, glo=True)
  rays['embed_idx'] = rays['embed_idx'] + 100            # table has 16 rows
  with pytest.raises(IndexError, match='embed_idx'):
    model.engine.check_embed_idx(rays['embed_idx'])
  # the kernels clamp / skip: no fault, finite results
  state, stats, _ = train_pstep(None, state, utils.Batch(rays=utils.Rays(**rays), rgb=gt), 0.1, None)
  torch.cuda.synchronize()
  assert np.isfinite(stats['loss']) and torch.isfinite(state.params).all()
  model.engine.close()
