"""Geodesic-polyhedron IPE basis (host constant).  Same contract as the reference's
MipNeRF360/internal/geopoly.py:78 `generate_basis` — including the *order* of the returned directions,
which fixes the IPE column order and therefore the layout of every trained first-layer weight.
tests/test_host_surface.py checks it against a fixture produced by the reference module itself.
"""
import itertools

import numpy as np


def _pairwise_sq_dist(a, b):
  """Squared distances between the columns of a [3,n] and b [3,m] (norm expansion, clamped at 0)."""
  na, nb = np.sum(a * a, axis=0), np.sum(b * b, axis=0)
  return np.maximum(na[:, None] + nb[None, :] - 2.0 * (a.T @ b), 0.0)


def _barycentric_grid(v):
  if v < 1:
    raise ValueError(f'v {v} must be >= 1')
  rows = [(i, j, v - i - j) for i in range(v + 1) for j in range(v + 1 - i)]
  return np.asarray(rows, dtype=np.float64) / v


def _tesselate(base_verts, base_faces, v, eps=1e-4):
  if not isinstance(v, int):
    raise ValueError(f'v {v} must an integer')
  w = _barycentric_grid(v)
  pts = []
  for face in base_faces:
    p = w @ base_verts[face, :]
    pts.append(p / np.sqrt(np.sum(p * p, axis=1, keepdims=True)))
  pts = np.concatenate(pts, axis=0)
  d = _pairwise_sq_dist(pts.T, pts.T)
  first = np.array([np.flatnonzero(row <= eps)[0] for row in d])   # first vertex each point coincides with
  return pts[np.unique(first), :]


def generate_basis(base_shape, angular_tesselation, remove_symmetries=True, eps=1e-4):
  """Returns the basis as an [n, 3] array (MLP.pos_basis_t is its transpose, models.py:395-396)."""
  if base_shape == 'icosahedron':
    phi = (np.sqrt(5) + 1) / 2
    verts = np.array([(-1, 0, phi), (1, 0, phi), (-1, 0, -phi), (1, 0, -phi), (0, phi, 1), (0, phi, -1),
                      (0, -phi, 1), (0, -phi, -1), (phi, 1, 0), (-phi, 1, 0), (phi, -1, 0),
                      (-phi, -1, 0)]) / np.sqrt(phi + 2)
    faces = np.array([(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8),
                      (5, 2, 3), (2, 7, 3), (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6),
                      (6, 1, 10), (9, 0, 11), (9, 11, 2), (9, 2, 5), (7, 2, 11)])
  elif base_shape == 'octahedron':
    verts = np.array([(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)], dtype=np.float64)
    corners = np.array(list(itertools.product([-1, 1], repeat=3)))
    pairs = np.argwhere(_pairwise_sq_dist(corners.T.astype(np.float64), verts.T) == 2)
    faces = np.sort(np.reshape(pairs[:, 1], [3, -1]).T, 1)
  else:
    raise ValueError(f'base_shape {base_shape} not supported')
  verts = _tesselate(verts, faces, angular_tesselation)
  if remove_symmetries:
    mirrored = _pairwise_sq_dist(verts.T, -verts.T) < eps
    verts = verts[np.any(np.triu(mirrored), axis=1), :]
  return verts[:, ::-1]
