"""Per-role cycle counters of the NerfMLP forward chain kernel (development instrumentation)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mipnerf360 as O
from tests import helpers as H
from nerf_hugs_b200.engine import Engine
from nerf_hugs_b200 import _lib
n = 4096
ocfg, ecfg = H.config_pair(precision='bf16_tc', max_rays=n)
params = O.init_params(ocfg, seed=0)
rays, gt = H.make_rays(n, seed=1)
eng = Engine(ecfg, H.basis_np())
flat = eng.flatten_params(params); eng.params_changed(flat)
fn = _lib.lib.hugs_debug_counters
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
for _ in range(3): eng.forward(flat, rays, 0.5, None, compute_extras=False, want_history=False)
torch.cuda.synchronize()
assert fn(eng._h, 1, None) == 0
eng.forward(flat, rays, 0.5, None, compute_extras=False, want_history=False)
buf = np.zeros((148, 16), np.int64)
assert fn(eng._h, 0, buf.ctypes.data_as(C.c_void_p)) == 0
rows = buf[buf[:, 4] > 0]
m = rows.mean(0)
tiles = 4096 / 148   # 128-sample tiles per SM (a CTA pair reports once for its two SMs)
print('reporting CTAs/clusters:', len(rows))
print('producer: total %.0f  wait_empty %.0f  wait_tile %.0f' % (m[0], m[1], m[2]))
print('mma     : total %.0f  wait_panel %.0f  wait_full %.0f  wait_feat/issue %.0f  (per tile: total %.0f panel %.0f full %.0f feat/issue %.0f)' % (m[4], m[5], m[6], m[7], m[4]/tiles, m[5]/tiles, m[6]/tiles, m[7]/tiles))
print('epi g0  : total %.0f  wait_acc %.0f  guard %.0f  work %.0f (per tile work %.0f)' % (m[8], m[9], m[10], m[8]-m[9]-m[10], (m[8]-m[9]-m[10])/tiles))
print('epi g3  : total %.0f  wait_acc %.0f  guard %.0f  work %.0f (per tile work %.0f)' % (m[12], m[13], m[14], m[12]-m[13]-m[14], (m[12]-m[13]-m[14])/tiles))
print('epi g0 detail (per tile): ld %.0f  math+store %.0f  publish %.0f' % (m[3]/tiles, m[11]/tiles, m[15]/tiles))
