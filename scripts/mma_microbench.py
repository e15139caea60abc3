"""tcgen05.mma rate micro-benchmark (one CTA, one issuing thread; see csrc/tc_microbench.cu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_hugs_b200 import _lib
fn = _lib.lib.hugs_debug_mma_rate
fn.restype = C.c_int; fn.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
torch.cuda.init(); torch.zeros(1, device='cuda')
out = (C.c_int64 * 2)()
print('N mode n_mmas | issue cyc/MMA | complete cyc/MMA')
for n in (64, 128, 256):
  for mode in (0, 1, 2, 3, 4, 6):
    for cnt in (64, 512):
      rc = fn(n, cnt, mode, out)
      assert rc == 0, _lib.lib.hugs_last_error()
      print(f'{n:4d} {mode:2d} {cnt:5d} | {out[0]/cnt:8.1f} | {out[1]/cnt:8.1f}')

f2 = _lib.lib.hugs_debug_mma_rate_cg2
f2.restype = C.c_int; f2.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
print('cta_group::2 (M=256): N mode n_mmas | issue cyc/MMA | complete cyc/MMA')
for n in (128, 256):
  for mode in (0, 1, 4, 5):
    for cnt in (64, 512):
      assert f2(n, cnt, mode, out) == 0, _lib.lib.hugs_last_error()
      print(f'{n:4d} {mode:2d} {cnt:5d} | {out[0]/cnt:8.1f} | {out[1]/cnt:8.1f}')

fl = _lib.lib.hugs_debug_ldtm_rate
fl.restype = C.c_int; fl.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
print('tcgen05.ld 32x32b.x32: warps cols | cycles per 128-lane x cols tile-read | bytes/cycle/SM')
for nw in (4, 8, 16):
  for cols in (64, 256):
    it = 200
    assert fl(nw, cols, it, out) == 0, _lib.lib.hugs_last_error()
    cyc = out[0] / it
    nbytes = nw * 32 * cols * 4
    print(f'{nw:3d} {cols:4d} | {cyc:9.1f} | {nbytes / cyc:8.1f}')
