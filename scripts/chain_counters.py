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
raw = np.zeros(148 * 16 + 74 * 60 + 8192, np.int64)
buf = raw[:148 * 16].reshape(148, 16)
assert fn(eng._h, 0, raw.ctypes.data_as(C.c_void_p)) == 0
rows = buf[buf[:, 4] > 0]
if len(rows) == 0:
  rows = buf[:74]
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
ext = raw[148 * 16:148 * 16 + 74 * 60].reshape(74, 60).astype(np.float64)
ext = ext[ext.sum(1) > 0]
print('epilogue latency (acc_full seen by group 0 -> panels ready at the issuer): %.0f cycles avg over %.0f waits' % (ext[:, 57].sum() / max(ext[:, 58].sum(), 1), ext[:, 58].mean()))
nw = max(ext[:, 58].sum(), 1)
print('  acc_full -> local group q arrived: ' + ' '.join('%.0f' % (ext[:, 53 + g].sum() / nw) for g in range(4)) + ' ; last local arrival -> issuer proceeds: %.0f' % (ext[:, 59].sum() / nw))
seg = raw[148 * 16:148 * 16 + 74 * 60].reshape(74, 20, 3).astype(np.float64)
seg = seg[seg.sum((1, 2)) > 0].mean(0) / units
print('per-segment waits per unit (panel, weights, features):')
for i in range(17):
  if seg[i].sum() > 0: print('  seg %2d: %7.0f %7.0f %7.0f' % (i, seg[i, 0], seg[i, 1], seg[i, 2]))
tr = raw[148 * 16 + 74 * 60:].reshape(4096, 2)
tr = tr[tr[:, 1] > 0]
if len(tr):
  tr = tr[np.argsort(tr[:, 1], kind='stable')]
  np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'trace.npy'), tr)
  print('trace events:', len(tr))
