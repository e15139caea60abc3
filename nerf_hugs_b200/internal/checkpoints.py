"""flax.training.checkpoints.{save,restore}_checkpoint for the TrainState of this package
(call sites: MipNeRF360/train.py:121,235,285, eval.py:74, render.py:122).

The reference stores `checkpoint_<step>` files: `flax.serialization.to_bytes(TrainState)`, i.e. msgpack of the state
dict {'step', 'params': {'params': {Module: {Dense_k: {kernel, bias}}}}, 'opt_state': {'0': {'count', 'mu', 'nu'},
'1': {'count'}}} (`optax.adam` = chain(scale_by_adam, scale_by_learning_rate(schedule)), train_utils.py:487-512) with
every array as msgpack ExtType 1 = packb((shape, dtype.name, bytes)) and arrays above 2**30 bytes split into chunks.

PARITY UNPINNED: flax / optax are not installable in the build container and the reference ships no checkpoint fixture,
so the encoding below restates flax.serialization's published format; tests cover round trips and a hand-assembled
byte string only.
"""
import os
import re
from typing import Any, Dict, Optional

import msgpack
import numpy as np
import torch

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_MAX_CHUNK = 2 ** 30
CHECKPOINT_RE = re.compile(r'^checkpoint_(\d+)$')


# ---- flax.serialization msgpack encoding -------------------------------------------------------------------
def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
  arr = np.asarray(arr)
  return msgpack.packb((arr.shape, arr.dtype.name, arr.tobytes('C')), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
  shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
  return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()


def _ext_pack(x):
  if isinstance(x, torch.Tensor):
    x = x.detach().cpu().numpy()
  if isinstance(x, np.ndarray):
    return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
  if isinstance(x, np.generic):
    return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
  if isinstance(x, complex):
    return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
  return x


def _ext_unpack(code, data):
  if code == _EXT_NDARRAY:
    return _ndarray_from_bytes(data)
  if code == _EXT_NPSCALAR:
    return _ndarray_from_bytes(data)[()]
  if code == _EXT_COMPLEX:
    re_, im = msgpack.unpackb(data)
    return complex(re_, im)
  return msgpack.ExtType(code, data)


def _chunk(tree):
  """Arrays above MAX_CHUNK bytes become {'__msgpack_chunked_array__', 'shape', 'chunks'} (flax.serialization._chunk)."""
  if isinstance(tree, dict):
    return {k: _chunk(v) for k, v in tree.items()}
  if isinstance(tree, torch.Tensor):
    tree = tree.detach().cpu().numpy()
  if isinstance(tree, np.ndarray) and tree.size * tree.dtype.itemsize > _MAX_CHUNK:
    flat = tree.reshape(-1)
    per = max(1, _MAX_CHUNK // tree.dtype.itemsize)
    chunks = {str(i): flat[o:o + per] for i, o in enumerate(range(0, flat.size, per))}
    return {'__msgpack_chunked_array__': True, 'shape': {str(i): int(s) for i, s in enumerate(tree.shape)}, 'chunks': chunks}
  return tree


def _unchunk(tree):
  if isinstance(tree, dict):
    if '__msgpack_chunked_array__' in tree:
      shape = tuple(tree['shape'][str(i)] for i in range(len(tree['shape'])))
      flat = np.concatenate([tree['chunks'][str(i)] for i in range(len(tree['chunks']))])
      return flat.reshape(shape)
    return {k: _unchunk(v) for k, v in tree.items()}
  return tree


def to_bytes(state_dict: Dict[str, Any]) -> bytes:
  """flax.serialization.msgpack_serialize of a state dict (nested dicts of arrays / scalars)."""
  return msgpack.packb(_chunk(state_dict), default=_ext_pack, strict_types=True)


def from_bytes(data: bytes) -> Dict[str, Any]:
  """flax.serialization.msgpack_restore."""
  return _unchunk(msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ---- TrainState <-> flax state dict ------------------------------------------------------------------------
def state_dict(state, model) -> Dict[str, Any]:
  """flax.serialization.to_state_dict(TrainState) of the reference (step, params, optax.adam chain state)."""
  eng = model.engine
  np_tree = lambda flat: _map(eng.unflatten_params(flat), lambda t: t.numpy().astype(np.float32))
  count = np.asarray(state.step, np.int32)
  return {'step': np.asarray(state.step, np.int32),
          'params': {'params': np_tree(state.params)},
          'opt_state': {'0': {'count': count, 'mu': {'params': np_tree(state.mu)}, 'nu': {'params': np_tree(state.nu)}},
                        '1': {'count': count}}}


def _map(tree, fn):
  return {k: _map(v, fn) for k, v in tree.items()} if isinstance(tree, dict) else fn(tree)


def load_state_dict(state, model, sd: Dict[str, Any]):
  """flax.serialization.from_state_dict: fills `state` (in place) from a reference-format state dict."""
  eng = model.engine
  params = sd['params'].get('params', sd['params'])
  state.params.copy_(eng.flatten_params(params))
  opt = sd.get('opt_state')
  if opt is not None and '0' in opt and 'mu' in opt['0']:
    mu, nu = opt['0']['mu'], opt['0']['nu']
    state.mu.copy_(eng.flatten_params(mu.get('params', mu)))
    state.nu.copy_(eng.flatten_params(nu.get('params', nu)))
  state.step = int(np.asarray(sd['step']))
  model._packed_version = None          # force the bf16 operand copies to be rebuilt
  return state


# ---- flax.training.checkpoints surface ---------------------------------------------------------------------
def latest_checkpoint(ckpt_dir: str, prefix: str = 'checkpoint_') -> Optional[str]:
  if not os.path.isdir(ckpt_dir):
    return None
  steps = [(int(m.group(1)), f) for f in os.listdir(ckpt_dir) for m in [CHECKPOINT_RE.match(f)] if m and f.startswith(prefix)]
  return os.path.join(ckpt_dir, max(steps)[1]) if steps else None


def save_checkpoint(ckpt_dir: str, target, step: int, model=None, keep: int = 1, overwrite: bool = False) -> str:
  """checkpoints.save_checkpoint(ckpt_dir, state, step, keep=...) (train.py:235-236): writes checkpoint_<step> and
  keeps the `keep` newest.  `target` is a TrainState (needs `model`) or an already-built state dict."""
  sd = target if isinstance(target, dict) else state_dict(target, model)
  os.makedirs(ckpt_dir, exist_ok=True)
  path = os.path.join(ckpt_dir, f'checkpoint_{int(step)}')
  if os.path.exists(path) and not overwrite:
    raise FileExistsError(f'{path} exists (flax raises InvalidCheckpointError here); pass overwrite=True')
  tmp = path + '.tmp'
  with open(tmp, 'wb') as f:
    f.write(to_bytes(sd))
  os.replace(tmp, path)
  olds = sorted((int(m.group(1)), f) for f in os.listdir(ckpt_dir) for m in [CHECKPOINT_RE.match(f)] if m)
  for _, f in olds[:-keep] if keep > 0 else []:
    os.remove(os.path.join(ckpt_dir, f))
  return path


def restore_checkpoint(ckpt_dir: str, target, model=None, step: Optional[int] = None):
  """checkpoints.restore_checkpoint(ckpt_dir, state) (train.py:121, eval.py:74, render.py:122): returns `target`
  unchanged when the directory holds no checkpoint (as flax does), else the restored state."""
  path = ckpt_dir if os.path.isfile(ckpt_dir) else (
      os.path.join(ckpt_dir, f'checkpoint_{int(step)}') if step is not None else latest_checkpoint(ckpt_dir))
  if path is None or not os.path.exists(path):
    return target
  with open(path, 'rb') as f:
    sd = from_bytes(f.read())
  if target is None:
    return sd
  return load_state_dict(target, model, sd)
