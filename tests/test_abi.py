"""The C-ABI shared library loads and exports every symbol include/hugs_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  src = open(os.path.join(ROOT, 'include', 'hugs_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(hugs_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  from nerf_hugs_b200 import _lib
  declared = _declared_symbols()
  assert len(declared) >= 14
  for name in declared:
    assert hasattr(_lib.lib, name), f'{name} declared in hugs_b200.h but not exported'
    assert name in _lib.SYMBOLS, f'{name} has no ctypes binding'
  assert _lib.lib.hugs_abi_version() == 2


def test_struct_sizes_match_header():
  from nerf_hugs_b200 import _lib
  # sizeof() of the C structs, printed by a g++ build of include/hugs_b200.h
  assert ctypes.sizeof(_lib.ModelDesc) == 536
  assert ctypes.sizeof(_lib.TensorDesc) == 88
  assert ctypes.sizeof(_lib.LevelOut) == 80
  assert ctypes.sizeof(_lib.Rays) == 72
  assert ctypes.sizeof(_lib.LossCfg) == 36
  assert ctypes.sizeof(_lib.AdamCfg) == 32
  assert ctypes.sizeof(_lib.CameraSet) == 112
  assert ctypes.sizeof(_lib.RayBatch) == 96
  assert ctypes.sizeof(_lib.NfRenderCfg) == 32
  assert ctypes.sizeof(_lib.TensorCopy) == 32
  assert ctypes.sizeof(_lib.FrameOut) == 48
  assert ctypes.sizeof(_lib.HashFieldDesc) == 64


def test_create_validates_and_fails_loudly_without_gpu():
  import torch
  from nerf_hugs_b200 import _lib
  d = _lib.ModelDesc()
  h = ctypes.c_void_p()
  rc = _lib.lib.hugs_create(ctypes.byref(d), ctypes.byref(h))
  assert rc == -1 and b'num_levels' in _lib.lib.hugs_last_error()
  if not torch.cuda.is_available():
    from nerf_hugs_b200.engine import Engine, EngineConfig
    import numpy as np
    with pytest.raises(RuntimeError, match='no CPU fallback'):
      Engine(EngineConfig(), np.zeros((3, 21), np.float32))
