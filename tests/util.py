"""Shared synthetic inputs for the tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def fcc(n, rho=0.8442, dtype=np.float32):
  """fcc lattice, n cells/side, N = 4 n^3; returns (R[N,3], L)."""
  a = (4.0 / rho) ** (1.0 / 3.0)
  basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
  g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n),
                           indexing='ij'), -1).reshape(-1, 1, 3)
  R = ((g + basis[None]) * a).reshape(-1, 3)
  return R.astype(dtype), np.float32(n * a)


def diamond(n, a=5.428, dtype=np.float64):
  """diamond-cubic lattice, 8 atoms/cell; returns (R, L)."""
  fccb = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
  basis = np.concatenate([fccb, fccb + 0.25])
  g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n),
                           indexing='ij'), -1).reshape(-1, 1, 3)
  R = ((g + basis[None]) * a).reshape(-1, 3)
  return R.astype(dtype), n * a


def jitter(R, L, scale, seed=0):
  rng = np.random.default_rng(seed)
  out = R + rng.normal(0, scale, R.shape).astype(R.dtype)
  return np.mod(out, R.dtype.type(L)).astype(R.dtype)


def momenta(N, dim, kT=1.0, seed=0, dtype=np.float32):
  rng = np.random.default_rng(seed)
  p = rng.normal(0, np.sqrt(kT), (N, dim))
  p = p - p.mean(axis=0, keepdims=True)
  return p.astype(dtype)


def sorted_rows(idx, N):
  return np.sort(idx, axis=-1)


def sparse_pairs(idx, N):
  """Sparse [2, cap] -> sorted array of (sender, receiver) pairs."""
  m = idx[0] < N
  p = np.stack([idx[1][m], idx[0][m]], 1).astype(np.int64)
  order = np.lexsort((p[:, 1], p[:, 0]))
  return p[order]
