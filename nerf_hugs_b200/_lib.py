"""ctypes binding of libhugs_b200.so (include/hugs_b200.h).  No torch types cross this boundary:
device pointers are passed as integers (tensor.data_ptr()), the stream as cudaStream_t.

There is no CPU fallback: if the shared library is missing this module raises at import.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HUGS_LIB selects another build of the same ABI (A/B timing of kernel variants during development)
LIB_PATH = os.environ.get('HUGS_LIB') or os.path.join(_HERE, 'libhugs_b200.so')


class HugsError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__(f'hugs_b200 error {code}: {msg}')
    self.code = code


class ModelDesc(C.Structure):
  _fields_ = [
      ('num_levels', C.c_int32), ('num_prop_samples', C.c_int32), ('num_nerf_samples', C.c_int32),
      ('nerf_depth', C.c_int32), ('nerf_width', C.c_int32), ('prop_depth', C.c_int32), ('prop_width', C.c_int32),
      ('bottleneck_width', C.c_int32), ('view_width', C.c_int32), ('skip_layer', C.c_int32),
      ('min_deg_point', C.c_int32), ('max_deg_point', C.c_int32), ('deg_view', C.c_int32),
      ('num_basis', C.c_int32), ('basis', C.c_float * 96),
      ('raydist_fn', C.c_int32), ('ray_shape', C.c_int32), ('nerf_contract', C.c_int32), ('prop_contract', C.c_int32),
      ('opaque_background', C.c_int32), ('bg_intensity', C.c_float),
      ('anneal_slope', C.c_float), ('dilation_multiplier', C.c_float), ('dilation_bias', C.c_float),
      ('resample_padding', C.c_float), ('near_anneal_rate', C.c_float), ('near_anneal_init', C.c_float),
      ('num_glo_features', C.c_int32), ('num_embeddings', C.c_int32),
      ('density_bias', C.c_float), ('rgb_premultiplier', C.c_float), ('rgb_bias', C.c_float), ('rgb_padding', C.c_float),
      ('precision', C.c_int32), ('max_rays', C.c_int32), ('encoding', C.c_int32), ('reserved_', C.c_int32 * 3),
  ]


class Rays(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('origins', 'directions', 'viewdirs', 'radii', 'near', 'far', 'lossmult',
                                        'static_mask', 'embed_idx')]


class LossCfg(C.Structure):
  _fields_ = [('data_loss_type', C.c_int32), ('charb_padding', C.c_float), ('data_loss_mult', C.c_float),
              ('data_coarse_loss_mult', C.c_float), ('interlevel_loss_mult', C.c_float),
              ('distortion_loss_mult', C.c_float), ('use_static_mask', C.c_int32),
              ('withmask_transient_weight', C.c_float), ('disable_multiscale_loss', C.c_int32)]


class AdamCfg(C.Structure):
  _fields_ = [('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
              ('grad_max_norm', C.c_float), ('grad_max_val', C.c_float), ('step', C.c_int32), ('grad_scale', C.c_float)]


class TensorDesc(C.Structure):
  _fields_ = [('name', C.c_char * 64), ('offset', C.c_int64), ('rows', C.c_int32), ('cols', C.c_int32),
              ('module', C.c_int32)]


class LevelOut(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('rgb', 'acc', 'distance_mean', 'distance_median', 'distance_p5',
                                        'distance_p95', 'sdist', 'weights', 'density', 'rgbs')]


class CameraSet(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('pixtocams', 'camtoworlds', 'heights', 'widths', 'pixel_offset', 'images',
                                        'images_u8', 'static_masks', 'nears', 'fars', 'embed_idxs')] + \
             [('near', C.c_float), ('far', C.c_float), ('distortion', C.c_void_p), ('camtype', C.c_int32),
              ('reserved_', C.c_int32)]


class NfRenderCfg(C.Structure):
  _fields_ = [('opaque_background', C.c_int32), ('density_activation', C.c_int32), ('density_bias', C.c_float),
              ('rgb_premultiplier', C.c_float), ('rgb_bias', C.c_float), ('rgb_padding', C.c_float),
              ('reserved_', C.c_int32 * 2)]


class TensorCopy(C.Structure):
  _fields_ = [('ptr', C.c_void_p), ('flat_off', C.c_int64), ('rows', C.c_int32), ('cols', C.c_int32),
              ('transpose', C.c_int32), ('ld', C.c_int32)]


class HashFieldDesc(C.Structure):
  _fields_ = [('n_levels', C.c_int32), ('features_per_level', C.c_int32), ('log2_hashmap_size', C.c_int32),
              ('base_res', C.c_int32), ('per_level_scale', C.c_float), ('hidden_dim', C.c_int32),
              ('geo_feat_dim', C.c_int32), ('hidden_dim_color', C.c_int32), ('appearance_dim', C.c_int32),
              ('num_embeddings', C.c_int32), ('bound', C.c_float), ('contract', C.c_int32), ('max_samples', C.c_int32),
              ('max_rays', C.c_int32), ('precision', C.c_int32), ('reserved_', C.c_int32)]


class FrameOut(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('rgb', 'acc', 'distance_mean', 'distance_median', 'rgb_u8', 'sse')]


class RayBatch(C.Structure):
  _fields_ = [(n, C.c_void_p) for n in ('origins', 'directions', 'viewdirs', 'radii', 'near', 'far', 'lossmult',
                                        'static_mask', 'embed_idx', 'cam_idx', 'pix_coords', 'rgb')]


# every symbol include/hugs_b200.h declares: (name, restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int32, C.c_float
SYMBOLS = {
    'hugs_last_error': (C.c_char_p, []),
    'hugs_abi_version': (C.c_int, []),
    'hugs_create': (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_P)]),
    'hugs_destroy': (C.c_int, [_P]),
    'hugs_param_count': (C.c_int64, [_P]),
    'hugs_param_layout': (C.c_int, [_P, C.POINTER(TensorDesc), _I, C.POINTER(_I)]),
    'hugs_params_changed': (C.c_int, [_P, _P, _P]),
    'hugs_sample_intervals': (C.c_int, [_P, _P, _P, _P, _F, _I, _I, _I, _F, _F, _P, _P, _P]),
    'hugs_invert_cdf': (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'hugs_max_dilate_weights': (C.c_int, [_P, _P, _I, _I, _F, _F, _F, _P, _P, _P]),
    'hugs_alpha_composite': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, C.POINTER(LevelOut), _P]),
    'hugs_ipe_features': (C.c_int, [_P, C.POINTER(Rays), _P, _I, _I, _I, _P, _P]),
    'hugs_debug_encode_bf16': (C.c_int, [_P, C.POINTER(Rays), _P, _I, _I, _I, _P, _P]),
    'hugs_forward': (C.c_int, [_P, _P, C.POINTER(Rays), _I, _F, _P, _I, _I, C.POINTER(LevelOut), _P]),
    'hugs_loss_and_grad': (C.c_int, [_P, _P, C.POINTER(Rays), _P, _I, _F, _P, C.POINTER(LossCfg), _P, _P, _P]),
    'hugs_set_train_rng': (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    'hugs_set_grad_ready_event': (C.c_int, [_P, _P]),
    'hugs_adam_step': (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(AdamCfg), _P, _P]),
    'hugs_adam_step_stats': (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(AdamCfg), _P, _P, _P]),
    'hugs_make_ray_batch': (C.c_int, [C.POINTER(CameraSet), _P, _P, _P, _I, C.POINTER(RayBatch), _P]),
    'hugs_render_frame': (C.c_int, [_P, _P, C.POINTER(CameraSet), _I, _I, _I, _I, _I, _F, _I, C.POINTER(FrameOut), _P]),
    'hugs_field_forward': (C.c_int, [_P, _P, C.POINTER(Rays), _P, _I, _I, _I, _I, _P, _P]),
    'hugs_field_backward': (C.c_int, [_P, _P, C.POINTER(Rays), _I, _I, _P, _P, _P]),
    'hugs_nf_sample_intervals': (C.c_int, [_P, _P, _P, _P, _I, _F, _F, _F, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P]),
    'hugs_nf_merge_bins': (C.c_int, [_P, _I, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P]),
    'hugs_nf_composite': (C.c_int, [C.POINTER(NfRenderCfg), _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    'hugs_nf_clip_depth': (C.c_int, [_P, _P, _I, _P]),
    'hugs_nf_composite_bwd': (C.c_int, [C.POINTER(NfRenderCfg), _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'hugs_nf_rgb_loss': (C.c_int, [_P, _P, _P, _F, _I, _F, _I, _P, _P, _P]),
    'hugs_nf_rgb_loss_bwd': (C.c_int, [_P, _P, _P, _F, _I, _P, _P]),
    'hugs_hashfield_create': (C.c_int, [C.POINTER(HashFieldDesc), C.POINTER(_P)]),
    'hugs_hashfield_destroy': (C.c_int, [_P]),
    'hugs_hashfield_grid_floats': (C.c_int64, [_P]),
    'hugs_hashfield_mlp_floats': (C.c_int64, [_P]),
    'hugs_hashfield_layout': (C.c_int, [_P, C.POINTER(TensorDesc), _I, C.POINTER(_I)]),
    'hugs_hashfield_level_info': (C.c_int, [_P, _I, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32)]),
    'hugs_hashfield_params_changed': (C.c_int, [_P, _P, _P]),
    'hugs_hashfield_encode': (C.c_int, [_P, _P, C.POINTER(Rays), _P, _I, _I, _P, _P]),
    'hugs_hashfield_forward': (C.c_int, [_P, _P, _P, C.POINTER(Rays), _P, _I, _I, _I, _I, _P, _P]),
    'hugs_hashfield_backward': (C.c_int, [_P, _P, _P, C.POINTER(Rays), _P, _I, _I, _P, _P, _P, _P]),
    'hugs_nf_distortion_loss': (C.c_int, [_P, _P, _I, _I, _P, _P, _P]),
    'hugs_nf_interlevel_loss': (C.c_int, [_P, _P, _I, _P, _P, _I, _I, _P, _P, _P]),
    'hugs_nf_scale': (C.c_int, [_P, _P, _F, C.c_int64, _P, _P]),
    'hugs_params_copy': (C.c_int, [_P, _I, _P, _I, _P, _P]),
    'hugs_launch_count': (C.c_int64, []),
    'hugs_profile_enable': (C.c_int, [_P, _I]),
    'hugs_profile_read': (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

if not os.path.exists(LIB_PATH):
  raise ImportError(
      f'{LIB_PATH} is missing: build it with `python -m nerf_hugs_b200.build` (nvcc, sm_100a). '
      'nerf_hugs_b200 has no CPU or PyTorch fallback path.')

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
  _fn = getattr(lib, _name)
  _fn.restype = _res
  _fn.argtypes = _args


KERNEL_CLASSES = ('sample', 'encode', 'chain_fwd_prop', 'chain_fwd_nerf', 'composite_loss', 'chain_bwd_nerf',
                  'chain_bwd_prop', 'wgrad_nerf', 'wgrad_prop', 'reductions', 'adam_pack', 'mlp_fp32')


def check(rc):
  if rc != 0:
    raise HugsError(rc, lib.hugs_last_error().decode('utf-8', 'replace'))
