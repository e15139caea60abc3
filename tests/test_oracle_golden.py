"""Pin the CPU oracle against outputs of the reference's own importable code.

Fixtures: tests/golden/*.npz, produced by tests/golden/make_golden.py from
/root/reference (nerfacto/utils/ray_utils.py, loss_utils.py, custom_functions.py,
MipNeRF360/internal/geopoly.py).  CPU only.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mipnerf360 as O

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_basis_golden_matches_reference_test_vector():
  """geopoly_test.py:76-100: the 21x3 icosahedral golden (as a set of rows)."""
  basis = np.load(f'{G}/geopoly_basis.npz')['icosahedron_2']
  assert basis.shape == (21, 3)
  golden_rows = np.array([[0.85065081, 0.0, 0.52573111], [0.80901699, 0.5, 0.30901699],
                          [0.52573111, 0.85065081, 0.0], [1.0, 0.0, 0.0],
                          [0.0, 0.0, 1.0], [-0.80901699, 0.5, -0.30901699]])
  for r in golden_rows:
    assert np.min(np.abs(basis - r).sum(-1)) < 1e-6


@pytest.mark.parametrize('name', ['small', 'prop', 'nerf', 'odd'])
@pytest.mark.parametrize('anneal', [1.0, 0.3])
def test_sample_intervals_matches_reference_torch_twin(name, anneal):
  """ray_utils.py:198-223 (torch twin of stepfun.py:214-263), deterministic branch."""
  z = np.load(f'{G}/sample_intervals.npz')
  t = torch.tensor(z[f'{name}_a{anneal}_t'])
  w = torch.tensor(z[f'{name}_a{anneal}_w'])
  ref = z[f'{name}_a{anneal}_out']
  # the reference's model code builds logits exactly like this (models.py:191-193)
  logits = torch.where(t[..., 1:] > t[..., :-1], anneal * torch.log(w),
                       torch.tensor(-float('inf')))
  out = O.sample_intervals(None, t, logits, ref.shape[-1] - 1, single_jitter=True, domain=(0., 1.))
  # the twin uses torch.linspace for u and searchsorted+gather; ours the brute-force
  # sorted_interp with float64-rounded linspace: agreement to a few ulps of the cdf.
  np.testing.assert_allclose(out.numpy(), ref, atol=2e-6, rtol=0)


def test_sample_single_interval_known_answer():
  """stepfun_test.py:579-586."""
  t = torch.tensor([1., 2, 3, 4, 5, 6])
  logits = torch.tensor([0., 0, 100, 0, 0])
  out = O.sample_intervals(None, t, logits, 10, single_jitter=True)
  np.testing.assert_allclose(out.numpy(), np.linspace(3, 4, 11), atol=1e-5, rtol=1e-5)
  ref = np.load(f'{G}/sample_intervals.npz')['single_out'][0]
  np.testing.assert_allclose(out.numpy(), ref, atol=1e-6)


def test_losses_match_reference_torch_twin():
  """loss_utils.py:7-31,65-77 vs stepfun.py:64-86,266-276."""
  z = np.load(f'{G}/losses.npz')
  t, w = torch.tensor(z['t']), torch.tensor(z['w'])
  te, we = torch.tensor(z['t_env']), torch.tensor(z['w_env'])
  np.testing.assert_allclose(O.lossfun_distortion(t, w).numpy(), z['distortion'], rtol=1e-5, atol=1e-7)
  _, outer = O.inner_outer(t, te, we)
  np.testing.assert_allclose(outer.numpy(), z['outer'], rtol=0, atol=2e-6)


def test_coord_matches_reference_torch_twin():
  """custom_functions.py:15-21,55-63 vs coord.py:21-27,136-147."""
  z = np.load(f'{G}/coord.npz')
  np.testing.assert_array_equal(O.contract(torch.tensor(z['x'])).numpy(), z['contract'])
  np.testing.assert_allclose(O.pos_enc(torch.tensor(z['v']), 0, 4, True).numpy(), z['pos_enc_0_4'],
                             atol=1e-6)
