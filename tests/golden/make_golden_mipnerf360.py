"""Golden fixtures from the reference's OWN Mip-NeRF 360 (JAX-path) source, executed on NumPy.

    python tests/golden/make_golden_mipnerf360.py          # in the build container (needs /root/reference)

JAX / flax / gin are not installable here, so `jax_numpy_shim.py` stands in for them (float32 NumPy, the published
jax definitions of interp / softmax / softplus / sigmoid, a dual-number `jax.linearize`).  The files below are loaded
from `/root/reference/MipNeRF360/internal/` unmodified and executed:

    math.py  stepfun.py  coord.py  render.py  geopoly.py  models.py (Model.__call__, MLP.__call__)
    train_utils.py: compute_data_loss / interlevel_loss / distortion_loss / tree_* / clip_gradients  (cut out by
                    name with `ast`, because the module itself imports optax / datasets / PIL-heavy code)

Outputs: tests/golden/mip360_ops.npz (operator level) and tests/golden/mip360_model.npz (whole Model.__call__ + the
three losses).  Network parameters are not stored: `tests/helpers.py::golden_params` regenerates them from a NumPy
seed (a checksum in the fixture guards against a drifting generator).
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import jax_numpy_shim as shim  # noqa: E402
from tests import helpers as H  # noqa: E402

REF = '/root/reference/MipNeRF360/internal'
F32 = np.float32


def load_reference():
  """Loads the reference modules under the package name `internal` with the shim installed."""
  stubs = ['internal.configs', 'internal.image', 'internal.utils', 'internal.camera_utils', 'internal.datasets']
  ctx = shim.installed(extra_stubs=stubs)
  ctx.__enter__()
  pkg = types.ModuleType('internal')
  pkg.__path__ = [REF]
  sys.modules['internal'] = pkg
  mods = {}
  for name in ('math', 'geopoly', 'stepfun', 'coord', 'render', 'models'):
    spec = importlib.util.spec_from_file_location(f'internal.{name}', f'{REF}/{name}.py')
    m = importlib.util.module_from_spec(spec)
    sys.modules[f'internal.{name}'] = m
    setattr(pkg, name, m)
    for s in stubs:
      setattr(pkg, s.split('.')[1], sys.modules[s])
    spec.loader.exec_module(m)
    mods[name] = m
  # functions of train_utils.py, cut out by name and executed against the same shim
  src = open(f'{REF}/train_utils.py').read()
  tree = ast.parse(src)
  wanted = {'compute_data_loss', 'interlevel_loss', 'distortion_loss', 'tree_sum', 'tree_norm_sq', 'tree_norm',
            'tree_abs_max', 'tree_len', 'clip_gradients'}
  body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
  assert {n.name for n in body} == wanted, {n.name for n in body} ^ wanted
  ns = {'jnp': shim.jnp, 'jax': shim.jax, 'stepfun': mods['stepfun'], 'collections': __import__('collections'),
        'np': np}
  exec(compile(ast.Module(body=body, type_ignores=[]), f'{REF}/train_utils.py', 'exec'), ns)
  mods['train_utils'] = types.SimpleNamespace(**{k: ns[k] for k in wanted})
  return mods


def ops_golden(R, out):
  rng = np.random.default_rng(2024)
  math, stepfun, coord, render = R['math'], R['stepfun'], R['coord'], R['render']

  # ---- math.safe_sin (quirk B12), coord.pos_enc, learning_rate_decay ----
  x = np.concatenate([rng.uniform(-5000, 5000, 512), [0., 314.15927, -314.15927, 314.2, -314.2, 4096.5, -4096.5,
                                                        100 * np.pi, 1e4, -1e4]]).astype(F32)
  out['safe_sin_x'] = x
  out['safe_sin_y'] = math.safe_sin(x)
  v = rng.normal(size=(32, 3)).astype(F32); v /= np.linalg.norm(v, axis=-1, keepdims=True)
  out['pos_enc_v'] = v
  out['pos_enc_out'] = coord.pos_enc(v, 0, 4, True)
  steps = np.array([0, 1, 100, 511, 512, 5000, 125000, 250000], np.int64)
  out['lr_steps'] = steps
  out['lr_values'] = np.array([float(math.learning_rate_decay(int(s), 2e-3, 2e-5, 250000, 512, 0.01)) for s in steps],
                              np.float64)

  # ---- math.sorted_interp ----
  xp = np.sort(rng.uniform(0, 1, (16, 33)).astype(F32), -1); xp[:, 0], xp[:, -1] = 0., 1.
  xp[:, 7] = xp[:, 6]                                   # a flat CDF segment
  fp = np.sort(rng.uniform(0, 1, (16, 33)).astype(F32), -1)
  xq = np.sort(rng.uniform(0, 1, (16, 40)).astype(F32), -1)
  xq[:, 3] = xp[:, 5]                                   # queries exactly on a knot
  out['si_x'], out['si_xp'], out['si_fp'] = xq, xp, fp
  out['si_out'] = math.sorted_interp(xq, xp, fp)

  # ---- coord.contract / track_linearize / lift_and_diagonalize / integrated_pos_enc ----
  basis = R['geopoly'].generate_basis('icosahedron', 2).T.astype(F32)      # [3, 21] == MLP.pos_basis_t
  n, S = 12, 16
  o = (rng.normal(size=(n, 3)) * np.array([1., 1., 1.])).astype(F32)
  d = rng.normal(size=(n, 3)).astype(F32); d /= np.linalg.norm(d, axis=-1, keepdims=True) / rng.uniform(0.8, 1.2, (n, 1))
  d = d.astype(F32)
  radii = rng.uniform(5e-4, 3e-3, (n, 1)).astype(F32)
  t = np.sort(np.concatenate([rng.uniform(0.2, 4.0, (n, S // 2 + 1)), rng.uniform(4.0, 3000.0, (n, S // 2))], -1), -1).astype(F32)
  out['cast_o'], out['cast_d'], out['cast_radii'], out['cast_t'] = o, d, radii, t
  for shape in ('cone', 'cylinder'):
    means, covs = render.cast_rays(t, o, d, radii, shape, diag=False)
    out[f'cast_{shape}_means'], out[f'cast_{shape}_covs'] = means, covs
  means, covs = render.cast_rays(t, o, d, radii, 'cone', diag=False)
  cm, cc = coord.track_linearize(coord.contract, means, covs)
  out['contract_means'], out['contract_covs'] = cm, cc
  lm, lv = coord.lift_and_diagonalize(cm, cc, basis)
  out['lift_means'], out['lift_vars'] = lm, lv
  out['ipe_contract'] = coord.integrated_pos_enc(lm, lv, 0, 12)            # |2^11 * mean| reaches ~4096 rad (B12)
  lm2, lv2 = coord.lift_and_diagonalize(means, covs, basis)                # no contraction (phototourism gins)
  near_mask = t[:, 1:] < 4.0                                               # keep arguments moderate without contraction
  out['ipe_plain_mask'] = near_mask
  out['ipe_plain'] = coord.integrated_pos_enc(lm2, lv2, 0, 12)

  # ---- coord.construct_ray_warps ----
  s = np.linspace(0, 1, 33).astype(F32)[None].repeat(4, 0)
  near = np.array([[0.2], [0.05], [1.0], [2.0]], F32); far = np.array([[1e6], [100.], [2.0], [6.0]], F32)
  out['warp_s'], out['warp_near'], out['warp_far'] = s, near, far
  fns = {'none': None, 'reciprocal': shim.jnp.reciprocal, 'log': shim.jnp.log, 'piecewise': 'piecewise'}
  for name, fn in fns.items():
    t_to_s, s_to_t = coord.construct_ray_warps(fn, near, far)
    tt = s_to_t(s)
    out[f'warp_{name}_t'] = tt
    out[f'warp_{name}_s_back'] = t_to_s(tt)

  # ---- stepfun: max_dilate_weights, sample_intervals, lossfun_outer / distortion, weighted_percentile ----
  nb = 64
  ts = np.sort(rng.uniform(0, 1, (24, nb + 1)).astype(F32), -1); ts[:, 0], ts[:, -1] = 0., 1.
  ts[3, 10] = ts[3, 9]                                                     # a zero-width interval
  w = (rng.uniform(0, 1, (24, nb)).astype(F32) ** 4); w /= w.sum(-1, keepdims=True)
  for dil in (0.0025 + 0.5 / 64, 0.02):
    td, wd = stepfun.max_dilate_weights(ts, w, F32(dil), domain=(0., 1.), renormalize=True)
    out[f'dilate_{dil:.4f}_t'], out[f'dilate_{dil:.4f}_w'] = td, wd
  out['dilate_in_t'], out['dilate_in_w'] = ts, w
  logits = np.where(ts[:, 1:] > ts[:, :-1], np.log(w), -np.inf).astype(F32)
  out['samp_logits'] = logits
  out['samp_det'] = stepfun.sample_intervals(None, ts, logits, 32, single_jitter=True, domain=(0., 1.))
  u01 = rng.uniform(0, 1, (24, 1)).astype(F32)
  key, _ = shim.jax.random.split(shim.KeyStream([u01]))
  out['samp_jitter_u'] = u01
  out['samp_jitter'] = stepfun.sample_intervals(key, ts, logits, 32, single_jitter=True, domain=(0., 1.))
  # integrate_weights / softmax CDF the samplers above went through (for the exact-index op test)
  wsm = shim.jax.nn.softmax(logits, axis=-1)
  out['samp_cdf'] = stepfun.integrate_weights(wsm)

  tq = np.sort(rng.uniform(0, 1, (24, 129)).astype(F32), -1)
  wq = rng.uniform(0, 1, (24, 128)).astype(F32); wq /= 1.3 * wq.sum(-1, keepdims=True)
  we = w / F32(1.1)
  out['loss_t'], out['loss_w'], out['loss_te'], out['loss_we'] = tq, wq, ts, we
  out['loss_outer'] = stepfun.lossfun_outer(tq, wq, ts, we)
  out['loss_distortion'] = stepfun.lossfun_distortion(tq, wq)
  out['inner_outer_in'], out['inner_outer_out'] = stepfun.inner_outer(tq, ts, we)[0], stepfun.inner_outer(tq, ts, we)[1]
  wp = rng.uniform(0, 1, (24, nb)).astype(F32); wp /= wp.sum(-1, keepdims=True)
  wp[5, 20:30] = 0; wp[5] /= wp[5].sum()                                   # a flat CDF stretch
  out['pct_w'] = wp
  out['pct_out'] = stepfun.weighted_percentile(ts, wp, [5, 50, 95])

  # ---- render.compute_alpha_weights / volumetric_rendering ----
  nr, S = 20, 48
  td = np.sort(rng.uniform(0.2, 30.0, (nr, S + 1)).astype(F32), -1)
  dens = (rng.uniform(0, 1, (nr, S)).astype(F32) ** 3) * 12
  dens[:2] = 0.                                                            # empty rays
  dens[2, :] = 1e4 * (np.arange(S) == 7)                                   # a delta
  dirs = rng.normal(size=(nr, 3)).astype(F32)
  rgbs = rng.uniform(size=(nr, S, 3)).astype(F32)
  tfar = np.full((nr, 1), 1e6, F32)
  out['vr_tdist'], out['vr_density'], out['vr_dirs'], out['vr_rgbs'], out['vr_far'] = td, dens, dirs, rgbs, tfar
  for opaque in (False, True):
    wts, alpha, trans = render.compute_alpha_weights(dens, td, dirs, opaque_background=opaque)
    out[f'vr_w_{int(opaque)}'], out[f'vr_alpha_{int(opaque)}'], out[f'vr_trans_{int(opaque)}'] = wts, alpha, trans
    r = render.volumetric_rendering(rgbs, wts, td, 1.0, tfar, True)
    for k, v in r.items():
      out[f'vr_{k}_{int(opaque)}'] = v
  # distance_mean with a non-positive midpoint and zero weight there (0 * -inf = nan -> nan_to_num quirk)
  td0 = td.copy(); td0[:, 0] = -td0[:, 1]                                  # t_mid[0] == 0
  dens0 = dens.copy(); dens0[:, 0] = 0.
  w0 = render.compute_alpha_weights(dens0, td0, dirs, opaque_background=False)[0]
  with np.errstate(all='ignore'):
    r0 = render.volumetric_rendering(rgbs, w0, td0, 1.0, tfar, True)
  out['vr_nan_tdist'], out['vr_nan_w'], out['vr_nan_distance_mean'] = td0, w0, r0['distance_mean']


def loss_cfg(**kw):
  base = dict(withmask_transient_weight=0.0, disable_multiscale_loss=False, data_loss_type='charb', charb_padding=0.001,
              data_coarse_loss_mult=0.0, data_loss_mult=1.0, interlevel_loss_mult=1.0, distortion_loss_mult=0.01,
              grad_max_val=0.0, grad_max_norm=0.001, transient_type=None, vis_num_rays=16)
  base.update(kw)
  return types.SimpleNamespace(**base)


def model_golden(R, out):
  models, tu = R['models'], R['train_utils']
  for case in H.GOLDEN_MODEL_CASES:
    c = H.GOLDEN_MODEL_CASES[case]
    rays, gt = H.make_rays(c['n'], seed=c['seed'], near=c['near'], far=c['far'])
    rays_np = {k: v.numpy() for k, v in rays.items()}
    tree = H.golden_params(c)
    out[f'{case}_param_checksum'] = np.array(H.param_checksum(tree), np.float64)
    cfg = loss_cfg(transient_type=c.get('transient'))
    raydist = {None: None, 'reciprocal': shim.jnp.reciprocal, 'log': shim.jnp.log, 'piecewise': 'piecewise'}[c['raydist']]
    warp = R['coord'].contract if c['contract'] else None
    # the gin bindings of the case (`NerfMLP.net_width = 256` ...)
    shim.gin.bindings = {
        'NerfMLP': dict(net_depth=c['nerf_depth'], net_width=c['width'], warp_fn=warp),
        'PropMLP': dict(net_depth=c['prop_depth'], net_width=c['width'], warp_fn=warp, disable_rgb=True)}
    model = models.Model(config=cfg, num_prop_samples=c['n_prop'], num_nerf_samples=c['n_nerf'],
                         num_levels=c['levels'], raydist_fn=raydist, opaque_background=c['opaque'],
                         num_glo_features=c['glo'], num_embeddings=16, ray_shape=c.get('ray_shape', 'cone'))
    Rays = types.SimpleNamespace
    rr = Rays(**rays_np)
    key = None
    if c['jitter']:
      draws = H.golden_jitter(c)
      stream = []
      for l in range(c['levels']):
        stream.append(draws[l])
      key = shim.KeyStream(stream)
    with np.errstate(all='ignore'):
      renderings, history = model.apply({'params': tree}, key, rr, c['train_frac'],
                                        True, False, False)
    for l in range(c['levels']):
      for k in ('rgb', 'acc', 'distance_mean', 'distance_median', 'distance_percentile_5', 'distance_percentile_95'):
        out[f'{case}_L{l}_{k}'] = renderings[l][k]
      out[f'{case}_L{l}_sdist'] = history[l]['sdist']
      out[f'{case}_L{l}_weights'] = history[l]['weights']
      out[f'{case}_L{l}_density'] = history[l]['density']
    out[f'{case}_Lf_rgbs'] = history[-1]['rgb']
    # the three losses (train_utils.py:413-448) on the same forward
    batch = types.SimpleNamespace(rgb=gt.numpy())
    ld, stats = tu.compute_data_loss(batch, rr, renderings, cfg, c.get('transient') == 'withmask')
    out[f'{case}_loss_data'] = np.asarray(ld['data'], np.float64)
    out[f'{case}_mses'] = np.asarray(stats['mses'], np.float64)
    out[f'{case}_loss_interlevel'] = np.asarray(tu.interlevel_loss(history, cfg), np.float64)
    out[f'{case}_loss_distortion'] = np.asarray(tu.distortion_loss(history, cfg), np.float64)
    shim.gin.bindings = {}

  # clip_gradients (train_utils.py:351-369) on a small random tree
  rng = np.random.default_rng(5)
  g = {'params': {'NerfMLP_0': {'Dense_0': {'kernel': rng.normal(size=(8, 4)).astype(F32) * 1e-3,
                                            'bias': rng.normal(size=(4,)).astype(F32) * 1e-3}},
                  'PropMLP_0': {'Dense_0': {'kernel': rng.normal(size=(8, 4)).astype(F32) * 1e-5,
                                            'bias': rng.normal(size=(4,)).astype(F32) * 1e-5}}}}
  clipped = tu.clip_gradients(g, loss_cfg(grad_max_val=0.002, grad_max_norm=0.001))
  for m in g['params']:
    for leaf in ('kernel', 'bias'):
      out[f'clip_in_{m}_{leaf}'] = g['params'][m]['Dense_0'][leaf]
      out[f'clip_out_{m}_{leaf}'] = clipped['params'][m]['Dense_0'][leaf]


def main():
  R = load_reference()
  ops, model = {}, {}
  ops_golden(R, ops)
  model_golden(R, model)
  np.savez_compressed(f'{HERE}/mip360_ops.npz', **{k: np.asarray(v) for k, v in ops.items()})
  np.savez_compressed(f'{HERE}/mip360_model.npz', **{k: np.asarray(v) for k, v in model.items()})
  for f in ('mip360_ops.npz', 'mip360_model.npz'):
    print(f, os.path.getsize(f'{HERE}/{f}') // 1024, 'KiB')


if __name__ == '__main__':
  main()
