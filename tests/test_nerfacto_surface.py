"""CPU tests of the torch twins' call surface (nerf_hugs_b200/nerfacto): names, shapes, initial weights and checkpoint
files equal the reference's (`/root/reference/nerfacto/models/nerf.py`, `utils/checkpoint_utils.py`), and the
arithmetic has no CPU path."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from tests import nerfacto_helpers as H


@pytest.fixture(scope='module')
def gold():
  return np.load(H.GOLDEN)


def test_registry_matches_reference_names():
  from nerf_hugs_b200.nerfacto.models import criterion_dict, model_config_dict, model_dict
  for d in (model_config_dict, model_dict, criterion_dict):
    assert 'nerf' in d      # models/__init__.py:4-17


@pytest.mark.parametrize('name', list(H.CASES))
def test_initial_weights_equal_the_reference(gold, name):
  # same seed, same creation order, same RNG consumption as the reference's constructors -> identical parameters
  case, model, _ = H.build(name)
  sd = model.state_dict()
  want = gold[f'{name}/weights_checksum']
  assert sum(v.numel() for v in sd.values()) == int(want[1])
  got = float(sum(v.double().abs().sum() for v in sd.values()))
  assert abs(got - want[0]) <= 1e-9 * abs(want[0])


def test_state_dict_names_and_param_groups(gold):
  _, model, _ = H.build('photo')
  names = set(model.state_dict().keys())
  ref = {k.split('/', 2)[2] for k in gold.files if k.startswith('photo/gsum/')}
  assert names == ref
  groups = model.get_params_dict()     # nerf.py:228-241
  assert set(groups) == {'field', 'appearance_embedding'}
  assert sum(p.numel() for p in groups['field']) + sum(p.numel() for p in groups['appearance_embedding']) == \
      sum(p.numel() for p in model.parameters())
  _, model, _ = H.build('cfg1')
  assert set(model.get_params_dict()) == {'field'}


def test_engine_layer_order():
  _, model, _ = H.build('cfg1')
  lin = model.field['fine'].linears()
  feat = 3 + 6 * 15
  assert [l.in_features for l in lin] == [feat, 256, 256, 256, 256, 256 + feat, 256, 256, 256, 256, 256 + 27, 128]
  assert [l.out_features for l in lin] == [256] * 8 + [1, 256, 128, 3]


def test_config_fields_cover_the_reference_yaml():
  from nerf_hugs_b200.nerfacto.models import model_config_dict
  fields = {f.name for f in dataclasses.fields(model_config_dict['nerf'])}
  # every key of the model sections of nerfacto/configs/*_nerf*.yml
  for k in ('net_width', 'max_deg_point', 'use_appearance_embedding', 'use_transient_embedding', 'appearance_embedding_dim',
            'transient_embedding_dim', 'eval_embedding', 'opaque_background', 'num_coarse_nerf_samples_per_ray',
            'num_fine_nerf_samples_per_ray', 'proposal_initial_sampler', 'rgb_loss_type', 'transient_type',
            'withmask_transient_weight'):
    assert k in fields


def test_out_of_scope_heads_fail_loudly():
  from nerf_hugs_b200.nerfacto.models import model_config_dict, model_dict
  C, M = model_config_dict['nerf'], model_dict['nerf']
  with pytest.raises(NotImplementedError):
    M(C(transient_type='nerfw', use_transient_embedding=True), 1.0, False, False)
  with pytest.raises(NotImplementedError):
    M(C(net_width=512), 1.0, False, False)
  with pytest.raises(ValueError):
    M(C(proposal_initial_sampler='log'), 1.0, False, False)


def test_no_cpu_path(gold):
  case, model, _ = H.build('cfg1')
  batch = H.load_batch(gold, 'cfg1')
  with pytest.raises(RuntimeError, match='CUDA'):
    model(batch=batch, curr_step=1, perturb=False)


def test_withmask_is_not_dispatched_like_the_reference():
  # quirk B4 (nerf.py:610-627): Loss.forward has no 'withmask' branch
  from nerf_hugs_b200.nerfacto.models import criterion_dict, model_config_dict, model_dict
  model = model_dict['nerf'](model_config_dict['nerf'](transient_type='withmask'), 1.0, False, False)
  with pytest.raises(NotImplementedError):
    criterion_dict['nerf'](model)({}, {}, (1, 1, 1), False, {})


def test_snapshot_file_layout(tmp_path):
  from nerf_hugs_b200.nerfacto.utils import checkpoint_utils as ck
  from nerf_hugs_b200.nerfacto.utils.utils import State
  _, model, _ = H.build('eval')
  opt = torch.optim.Adam([{'params': v} for v in model.get_params_dict().values()], lr=1e-3)
  f = str(tmp_path / 'checkpoint_00000007.ckpt')
  ck.save_snapshot(f, State(step=7, epoch=2, next_eval_idx=1), model, opt, None, None)
  raw = torch.load(f)
  assert set(raw) == {'state', 'model', 'optimizer', 'scheduler', 'scaler'}      # checkpoint_utils.py:33-39
  assert raw['state'] == {'step': 7, 'epoch': 2, 'next_eval_idx': 1}
  _, other, _ = H.build('eval')
  with torch.no_grad():
    for p in other.parameters():
      p.zero_()
  st = ck.load_snapshot(f, other, None, None, None, 'cpu')
  assert st.step == 7 and st.epoch == 2
  for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
    assert torch.equal(a, b), k
  f2 = str(tmp_path / 'w.ckpt')
  ck.save_weights(f2, State(step=3), model)
  assert ck.load_weights(f2, other, 'cpu').step == 3


def test_split_merge_tensor_data():
  from nerf_hugs_b200.nerfacto.utils.utils import merge_tensor_data, split_tensor_data
  d = {'a': torch.arange(10.).reshape(10, 1), 'b': [torch.arange(20.).reshape(10, 2), torch.arange(10)]}
  parts = split_tensor_data(d, 4)
  assert len(parts) == 3 and parts[2]['a'].shape[0] == 2 and parts[0]['b'][0].shape == (4, 2)
  back = merge_tensor_data(parts)
  assert torch.equal(back['a'], d['a']) and torch.equal(back['b'][1], d['b'][1])


# ------------------------------------------------------------------------------------------------ nerfacto (hash grid)
@pytest.fixture(scope='module')
def gold_hash():
  return np.load(H.GOLDEN_HASH)


@pytest.mark.parametrize('name', list(H.HASH_CASES))
def test_nerfacto_state_dict_equals_the_reference(gold_hash, name):
  # same module tree (buffers, `mlp_base.0.params`, `direction_encoder.params`, Linear indices), same seed -> same values
  case, model, _ = H.build_hash(name)
  sd = model.state_dict()
  assert sorted(sd.keys()) == list(gold_hash[f'{name}/state_keys'])
  want = gold_hash[f'{name}/weights_checksum']
  assert sum(v.numel() for v in sd.values()) == int(want[1])
  got = float(sum(v.double().abs().sum() for v in sd.values()))
  assert abs(got - want[0]) <= 1e-9 * abs(want[0])


def test_nerfacto_param_groups_and_loud_failures():
  from nerf_hugs_b200.nerfacto.models import model_config_dict, model_dict
  case, model, _ = H.build_hash('withmask')
  assert set(model.get_params_dict()) == {'field', 'proposal', 'appearance_embedding'}      # nerfacto.py:250-264
  C, M = model_config_dict['nerfacto'], model_dict['nerfacto']
  base = dict(H.HASH_CASES['contract']['model'])
  with pytest.raises(NotImplementedError):
    M(C(**{**base, 'enable_tcnn_mlp': True}), 2.0, True, False)
  with pytest.raises(NotImplementedError):
    M(C(**{**base, 'hidden_dim': 64}), 2.0, False, False)
  with pytest.raises(NotImplementedError):
    M(C(**{**base, 'transient_type': 'robustnerf'}), 2.0, False, False)
  with pytest.raises(AssertionError):
    M(C(**base), 1.0, False, True)       # nerfacto.py:133: contraction needs bound == 2


def test_nerfacto_no_cpu_path(gold_hash):
  case, model, _ = H.build_hash('eval')
  model.eval()
  with pytest.raises(RuntimeError, match='CUDA'):
    model(batch=H.load_hash_batch(gold_hash, 'eval'), curr_step=1, perturb=False, chunk_size=32)


def test_hashgrid_oracle_level_table():
  # the restated tcnn geometry: phototourism_nerfacto_withmask.yml's 16 levels, 2^21 entries, base 16 -> 8192
  from oracle import hashgrid as hg
  growth = np.exp((np.log(8192) - np.log(16)) / 15)
  levels, total = hg.level_table(16, 16, growth, 21)
  assert levels[0][1] == 16 and levels[-1][1] == 8192
  assert levels[0][3] == 4096 and all(cnt % 8 == 0 for _, _, _, cnt in levels)
  assert total * 2 == 47857600
  # dense levels index without hashing: corner (x, y, z) -> x + y * res + z * res^2
  x, y, z = torch.tensor([3]), torch.tensor([5]), torch.tensor([7])
  assert int(hg.grid_index(x, y, z, 16, 4096)) == 3 + 5 * 16 + 7 * 256
  assert int(hg.grid_index(x, y, z, 8192, 1 << 21)) == ((3 * 1) ^ ((5 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)) % (1 << 21)
