"""Snapshot files of the torch twins (`/root/reference/nerfacto/utils/checkpoint_utils.py:9-65`): one `torch.save`d
dict {'state': asdict(State), 'model': state_dict, 'optimizer', 'scheduler', 'scaler'} - the same keys, and (because the
parameter containers carry the reference's module names) the same state_dict, so the files are interchangeable."""
from dataclasses import asdict

import torch
from torch.nn.parallel import DataParallel as DP
from torch.nn.parallel import DistributedDataParallel as DDP

from .utils import State


def _unwrap(model):
  return model.module if isinstance(model, (DDP, DP)) else model


def load_snapshot(ckpt_file: str, model=None, optimizer=None, scheduler=None, scaler=None, device: str = 'cpu') -> State:
  ckpt_dict = torch.load(ckpt_file, map_location={'cuda:0': device})
  for obj, key in ((model, 'model'), (optimizer, 'optimizer'), (scheduler, 'scheduler'), (scaler, 'scaler')):
    if obj is not None:
      (_unwrap(obj) if key == 'model' else obj).load_state_dict(ckpt_dict[key])
  return State(**ckpt_dict['state'])


def save_snapshot(ckpt_file: str, state: State, model=None, optimizer=None, scheduler=None, scaler=None):
  state_dict = {'state': asdict(state), 'model': None, 'optimizer': None, 'scheduler': None, 'scaler': None}
  if model is not None:
    state_dict['model'] = _unwrap(model).state_dict()
  for obj, key in ((optimizer, 'optimizer'), (scheduler, 'scheduler'), (scaler, 'scaler')):
    if obj is not None:
      state_dict[key] = obj.state_dict()
  torch.save(state_dict, ckpt_file)


def load_weights(ckpt_file: str, model, device: str) -> State:
  ckpt_dict = torch.load(ckpt_file, map_location={'cuda:0': device})
  _unwrap(model).load_state_dict(ckpt_dict['model'])
  return State(**ckpt_dict['state'])


def save_weights(ckpt_file: str, state: State, model):
  torch.save({'state': asdict(state), 'model': _unwrap(model).state_dict()}, ckpt_file)
