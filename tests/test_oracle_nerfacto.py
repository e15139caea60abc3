"""oracle/nerfacto.py against outputs of the reference's own torch functions (tests/golden/nerfacto_ops.npz, generated
by tests/golden/make_golden.py from /root/reference/nerfacto/utils/ray_utils.py).  Groundwork for SURVEY.md §8f item 1."""
import os

import numpy as np
import pytest

from oracle import nerfacto as ON

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nerfacto_ops.npz'))


@pytest.mark.parametrize('opaque', [0, 1])
def test_density_to_weight_and_renders_match_reference(opaque):
  w, a, t = ON.density_to_weight(G['dens'], G['bins'], G['dirs'], bool(opaque))
  np.testing.assert_allclose(w, G[f'w_{opaque}'], rtol=2e-6, atol=3e-7)
  np.testing.assert_allclose(a, G[f'a_{opaque}'], rtol=2e-6, atol=3e-7)
  np.testing.assert_allclose(t, G[f't_{opaque}'], rtol=2e-6, atol=3e-7)
  np.testing.assert_allclose(ON.render_features(G[f'w_{opaque}'], G['feats'], G['bg']), G[f'rgb_{opaque}'], rtol=2e-6, atol=1e-6)
  np.testing.assert_allclose(ON.render_depth(G[f'w_{opaque}'], G['bins']), G[f'depth_{opaque}'], rtol=2e-6, atol=1e-6)


def test_quirk_b2_deltas_are_measured_from_the_first_fencepost():
  """The torch compositor is NOT the JAX one (render.py:132 uses adjacent deltas): keep them apart."""
  w_ref = G['w_0']
  adj = np.diff(G['bins'], axis=-1) * np.linalg.norm(G['dirs'], axis=-1, keepdims=True)
  dd = G['dens'] * adj
  w_adj = (1 - np.exp(-dd)) * np.exp(-np.concatenate([np.zeros_like(dd[:, :1]), np.cumsum(dd[:, :-1], -1)], -1))
  assert np.abs(w_adj - w_ref).max() > 1e-2


def test_pdf_and_uniform_sampling_match_reference():
  # a one-ulp difference of the CDF moves a sample by ulp / pdf of its bin: almost all samples agree to 2e-6, all to 1e-4
  err = np.abs(ON.pdf_sample(G['pdf_bins'], G['pdf_w'], 32) - G['pdf_out'])
  assert (err <= 2e-6).mean() > 0.995 and err.max() < 1e-4, (float((err <= 2e-6).mean()), float(err.max()))
  np.testing.assert_array_equal(ON.uniform_sample(5, 16), G['uniform_out'])
