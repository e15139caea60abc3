"""Host helpers of the torch twins (`/root/reference/nerfacto/utils/utils.py:11-107`): the training-state record and
the chunking of a batch dict for whole-image rendering."""
from dataclasses import dataclass

import torch
from torch import Tensor


@dataclass
class State:
  step: int = 0
  epoch: int = 0
  next_eval_idx: int = 0


def split_tensor_data(data, chunk_size: int) -> list:
  """Tensor / list / dict of tensors -> list of the same structure, split along dim 0 (utils.py:59-82)."""
  if isinstance(data, Tensor):
    return list(torch.split(data, split_size_or_sections=chunk_size, dim=0))
  if isinstance(data, (list, dict)):
    items = list(enumerate(data)) if isinstance(data, list) else list(data.items())
    parts = {k: split_tensor_data(v, chunk_size) for k, v in items}
    n = len(next(iter(parts.values())))
    if isinstance(data, list):
      return [[parts[k][j] for k, _ in items] for j in range(n)]
    return [{k: parts[k][j] for k, _ in items} for j in range(n)]
  raise NotImplementedError()


def merge_tensor_data(datas: list):
  """Inverse of split_tensor_data (utils.py:85-107)."""
  first = datas[0]
  if isinstance(first, Tensor):
    return torch.cat(datas, dim=0)
  if isinstance(first, list):
    return [merge_tensor_data([d[j] for d in datas]) for j in range(len(first))]
  if isinstance(first, dict):
    return {k: merge_tensor_data([d[k] for d in datas]) for k in first.keys()}
  raise NotImplementedError()
