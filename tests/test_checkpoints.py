"""flax-format checkpoint encoding (nerf_hugs_b200/internal/checkpoints.py): msgpack ext types, chunking, naming."""
import os

import msgpack
import numpy as np

from nerf_hugs_b200.internal import checkpoints as ck


def _tree(rng):
  dense = lambda i, o: {'kernel': rng.normal(size=(i, o)).astype(np.float32), 'bias': rng.normal(size=(o,)).astype(np.float32)}
  return {'NerfMLP_0': {'Dense_0': dense(5, 4), 'Dense_1': dense(4, 3)}, 'PropMLP_0': {'Dense_0': dense(5, 2)}}


def test_round_trip_preserves_structure_dtypes_and_values(tmp_path):
  rng = np.random.default_rng(0)
  sd = {'step': np.asarray(25000, np.int32), 'params': {'params': _tree(rng)},
        'opt_state': {'0': {'count': np.asarray(25000, np.int32), 'mu': {'params': _tree(rng)}, 'nu': {'params': _tree(rng)}},
                      '1': {'count': np.asarray(25000, np.int32)}}}
  path = ck.save_checkpoint(str(tmp_path), sd, 25000, keep=2)
  assert os.path.basename(path) == 'checkpoint_25000'           # flax.training.checkpoints naming
  back = ck.restore_checkpoint(str(tmp_path), None)
  assert int(back['step']) == 25000 and back['step'].dtype == np.int32
  a, b = sd['params']['params']['NerfMLP_0']['Dense_1'], back['params']['params']['NerfMLP_0']['Dense_1']
  assert b['kernel'].dtype == np.float32 and np.array_equal(a['kernel'], b['kernel']) and np.array_equal(a['bias'], b['bias'])
  assert sorted(back['opt_state']) == ['0', '1'] and sorted(back['opt_state']['0']) == ['count', 'mu', 'nu']
  # keep=2: a third save drops the oldest; restore picks the newest
  ck.save_checkpoint(str(tmp_path), sd, 26000, keep=2)
  ck.save_checkpoint(str(tmp_path), sd, 27000, keep=2)
  assert sorted(os.listdir(tmp_path)) == ['checkpoint_26000', 'checkpoint_27000']
  assert os.path.basename(ck.latest_checkpoint(str(tmp_path))) == 'checkpoint_27000'
  assert ck.restore_checkpoint(str(tmp_path / 'missing'), 'unchanged') == 'unchanged'     # flax returns the target


def test_decodes_a_hand_assembled_flax_byte_string():
  """ExtType 1 = packb((shape, dtype.name, C-order bytes)), ExtType 3 = numpy scalar (flax.serialization)."""
  arr = np.arange(6, dtype=np.float32).reshape(2, 3)
  ext = lambda a: msgpack.ExtType(1, msgpack.packb((a.shape, a.dtype.name, a.tobytes()), use_bin_type=True))
  scal = msgpack.ExtType(3, msgpack.packb(((), 'int32', np.int32(7).tobytes()), use_bin_type=True))
  blob = msgpack.packb({'step': scal, 'params': {'params': {'M': {'kernel': ext(arr)}}}}, use_bin_type=True)
  sd = ck.from_bytes(blob)
  assert sd['step'] == 7 and np.array_equal(sd['params']['params']['M']['kernel'], arr)
  # and our writer emits exactly that encoding for arrays
  again = msgpack.unpackb(ck.to_bytes({'k': arr}), raw=False)
  assert isinstance(again['k'], msgpack.ExtType) and again['k'].code == 1
  shape, name, buf = msgpack.unpackb(again['k'].data, raw=False)
  assert tuple(shape) == (2, 3) and name == 'float32' and buf == arr.tobytes()


def test_chunked_arrays_follow_the_flax_layout(monkeypatch):
  monkeypatch.setattr(ck, '_MAX_CHUNK', 64)
  arr = np.arange(100, dtype=np.float32).reshape(4, 25)
  raw = msgpack.unpackb(ck.to_bytes({'w': arr}), ext_hook=ck._ext_unpack, raw=False)
  assert raw['w']['__msgpack_chunked_array__'] is True and raw['w']['shape'] == {'0': 4, '1': 25}
  assert len(raw['w']['chunks']) == 7 and raw['w']['chunks']['0'].size == 16
  assert np.array_equal(ck.from_bytes(ck.to_bytes({'w': arr}))['w'], arr)
