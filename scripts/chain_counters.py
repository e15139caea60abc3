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
units = 4096 / len(rows) / (4 if len(rows) <= 74 else 2)   # program passes per reporting CTA (pair)
print('per unit (one pass of the segment program):')
print('mma  : total %.0f  wait_panel %.0f  wait_full %.0f  wait_feat %.0f  issue %.0f' % (m[4]/units, m[5]/units, m[6]/units, m[7]/units, (m[4]-m[5]-m[6]-m[7])/units))
for nm, o in (('epi g0', 8), ('epi g3', 12)):
  print('%s: total %.0f  wait_acc %.0f  work(incl publish) %.0f  publish %.0f' % (nm, m[o]/units, m[o+1]/units, m[o+2]/units, m[o+3]/units))
print('epi g0 relu/linear layers: tcgen05.ld+wait %.0f  cvt+st.shared %.0f' % (m[0]/units, m[1]/units))
print('epi g0 by type: view layers %.0f  density-head layer %.0f' % (m[2]/units, m[3]/units))
