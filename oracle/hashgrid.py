"""CPU restatement (torch) of the tiny-cuda-nn encodings the reference's nerfacto path calls (TEST INFRASTRUCTURE ONLY).

tiny-cuda-nn is a third-party dependency that is absent from /root/reference (requirements_torch.txt:8, unpinned git
HEAD); call sites: nerfacto/models/nerfacto.py:693-700 (SphericalHarmonics, degree 4), :716-772 / :923-954 (HashGrid).
This file restates its published algorithm (Mueller et al. 2022, "Instant Neural Graphics Primitives", and the public
`grid.h` / `spherical_harmonics.h` of the library):

  * level l: scale = exp2(l * log2(per_level_scale)) * base_resolution - 1 (float32), resolution = ceil(scale) + 1,
    entries = min(round_up(resolution^3, 8), 2^log2_hashmap_size); levels back to back, features innermost;
  * pos = x * scale + 0.5, cell = floor(pos), w = pos - cell; corner index = x + y * res + z * res^2 while the stride
    still fits the level's entry count, otherwise x ^ y * 2654435761 ^ z * 805459861 (uint32), modulo the entry count;
  * trilinear interpolation; parameters initialised uniformly in [-1e-4, 1e-4];
  * SphericalHarmonics(degree 4) evaluates the 16 real harmonics of 2 * input - 1.

PARITY UNPINNED against tcnn itself (no tcnn build, no golden vectors anywhere in the reference); pinned instead through
the reference's own call sites: tests/golden/make_golden_nerfacto.py runs the unmodified nerfacto.py on these modules.
`Encoding` mirrors the constructor / attributes of `tinycudann.Encoding` that nerfacto.py uses, so that this module can
stand in for `tinycudann` when the reference is imported.
"""
import math

import numpy as np
import torch
import torch.nn as nn

PRIMES = (1, 2654435761, 805459861)
MASK32 = 0xFFFFFFFF


def level_table(n_levels, base_resolution, per_level_scale, log2_hashmap_size):
  """-> list of (scale float32, resolution, offset, entries) per level, float32 arithmetic as in tcnn."""
  out, off = [], 0
  l2s = np.log2(np.float32(per_level_scale)).astype(np.float32)
  for l in range(n_levels):
    scale = np.float32(np.exp2(np.float32(l) * l2s).astype(np.float32) * np.float32(base_resolution) - np.float32(1.0))
    res = int(np.ceil(scale)) + 1
    cnt = min(res ** 3, MASK32 // 2)
    cnt = (cnt + 7) // 8 * 8
    cnt = min(cnt, 1 << log2_hashmap_size)
    out.append((scale, res, off, cnt))
    off += cnt
  return out, off


def grid_index(x, y, z, res, entries):
  """uint32 arithmetic on int64 tensors."""
  stride, index = 1, x.clone()
  stride *= res
  if stride <= entries:
    index = (index + y * stride) & MASK32
    stride *= res
  if stride <= entries:
    index = (index + z * stride) & MASK32
    stride *= res
  if entries < stride:
    index = ((x * PRIMES[0]) & MASK32) ^ ((y * PRIMES[1]) & MASK32) ^ ((z * PRIMES[2]) & MASK32)
  return index % entries


def hashgrid_encode(x, params, levels, features_per_level=2):
  """x [N, 3] float32 in [0, 1]; params flat [total_entries * F] -> [N, n_levels * F] (level-major)."""
  F = features_per_level
  table = params.view(-1, F)
  outs = []
  for scale, res, off, cnt in levels:
    # fmaf(scale, x, 0.5): one rounding (the product of two float32 numbers is exact in float64)
    pos = (x.double() * float(scale) + 0.5).to(x.dtype)
    cell = torch.floor(pos)
    w = pos - cell
    c = cell.to(torch.int64)
    acc = 0
    for k in range(8):
      d = [(k >> i) & 1 for i in range(3)]
      wk = 1
      for i in range(3):
        wk = wk * (w[:, i] if d[i] else 1 - w[:, i])
      idx = grid_index((c[:, 0] + d[0]) & MASK32, (c[:, 1] + d[1]) & MASK32, (c[:, 2] + d[2]) & MASK32, res, cnt) + off
      acc = acc + wk[:, None] * table[idx]
    outs.append(acc)
  return torch.cat(outs, -1)


def sh4(v01):
  """tcnn SphericalHarmonics(degree 4) of v01 in [0, 1]^3 -> [N, 16]."""
  v = v01 * 2 - 1
  x, y, z = v[:, 0], v[:, 1], v[:, 2]
  xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
  return torch.stack([
      torch.full_like(x, 0.28209479177387814),
      -0.48860251190291987 * y,
      0.48860251190291987 * z,
      -0.48860251190291987 * x,
      1.0925484305920792 * xy,
      -1.0925484305920792 * yz,
      0.94617469575755997 * z2 - 0.31539156525251999,
      -1.0925484305920792 * xz,
      0.54627421529603959 * x2 - 0.54627421529603959 * y2,
      0.59004358992664352 * y * (-3.0 * x2 + y2),
      2.8906114426405538 * xy * z,
      0.45704579946446572 * y * (1.0 - 5.0 * z2),
      0.3731763325901154 * z * (5.0 * z2 - 3.0),
      0.45704579946446572 * x * (1.0 - 5.0 * z2),
      1.4453057213202769 * z * (x2 - y2),
      0.59004358992664352 * x * (-x2 + 3.0 * y2),
  ], -1)


class Encoding(nn.Module):
  """Stand-in for `tinycudann.Encoding` (constructor and attributes as used by nerfacto.py)."""

  def __init__(self, n_input_dims, encoding_config, dtype=None, seed=1337):
    super().__init__()
    self.n_input_dims = n_input_dims
    self.encoding_config = dict(encoding_config)
    self.otype = encoding_config['otype']
    if self.otype == 'HashGrid':
      assert n_input_dims == 3
      self.F = int(encoding_config['n_features_per_level'])
      self.levels, total = level_table(int(encoding_config['n_levels']), int(encoding_config['base_resolution']),
                                       float(encoding_config['per_level_scale']), int(encoding_config['log2_hashmap_size']))
      self.n_output_dims = len(self.levels) * self.F
      self.params = nn.Parameter((torch.rand(total * self.F, dtype=torch.float32) * 2 - 1) * 1e-4)
    elif self.otype == 'SphericalHarmonics':
      assert n_input_dims == 3 and int(encoding_config['degree']) == 4
      self.n_output_dims = 16
      self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32))
    else:
      raise NotImplementedError(self.otype)

  def forward(self, x):
    x = x.to(torch.float32)
    if self.otype == 'HashGrid':
      return hashgrid_encode(x, self.params, self.levels, self.F)
    return sh4(x)
