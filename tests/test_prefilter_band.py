"""CPU check of the stencil scan's pre-filter bound (csrc/jmd_neighbor.cu: fill()).

The kernel accepts a candidate when its contracted, image-shifted a^2 is below
cutoff^2 - band, rejects it above cutoff^2 + band and evaluates the reference's
exact op sequence only in between.  That is safe iff |a^2 - d^2_exact| < band for
both orientations of the exact test whenever both atoms are "regular".  Here the two
computations are emulated in NumPy float32 / float64 on millions of pairs placed
within a few ulps of the cutoff (and anywhere else), for the box sizes of the
BASELINE configs."""
import numpy as np
import pytest


def _band(L, cutoff_sq, dim, dtype):
  u = 5.9604644775390625e-08 if dtype == np.float32 else 1.1102230246251565e-16
  c = np.sqrt(cutoff_sq)
  delta = 10.0 * u * max(L, c)
  band = dim * (2.0 * (c + delta) + delta) * delta + 2.0 * (dim + 1) * u * 1.01 * cutoff_sq
  return 2.0 * band          # P.band; f_lo / f_hi = cutoff_sq -/+ P.band


def _exact_d2(a, b, L, dtype):
  """space.py:213-235 in separately rounded ops: mod(fl(a-b) + L/2, L) - L/2, sum of squares."""
  L = dtype(L)
  h = dtype(L * dtype(0.5))
  d = (a - b).astype(dtype)
  t = (d + h).astype(dtype)
  m = np.mod(t, L).astype(dtype)
  r = (m - h).astype(dtype)
  sq = (r * r).astype(dtype)
  acc = sq[:, 0]
  for k in range(1, sq.shape[1]):
    acc = (acc + sq[:, k]).astype(dtype)
  return acc


def _filter_a2(a, b, shift, dtype):
  """hs = fl(a + shift); a2 = fma chain of (hs - b): emulated with one rounding per fma."""
  hs = (a + shift.astype(dtype)).astype(dtype)
  x = (hs - b).astype(dtype).astype(np.float64 if dtype == np.float32 else np.longdouble)
  acc = (x[:, 0] * x[:, 0]).astype(dtype)
  for k in range(1, x.shape[1]):
    acc = (x[:, k] * x[:, k] + acc.astype(x.dtype)).astype(dtype)
  return acc


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('L,cells', [(33.592, 11), (105.81, 37), (335.92, 119)])
def test_prefilter_error_is_inside_the_band(dtype, L, cells):
  rng = np.random.default_rng(int(L))
  dim, cut = 3, 2.8
  cutoff_sq = float(dtype(cut) ** 2)
  band = _band(L, cutoff_sq, dim, dtype)
  n = 400_000
  cs = L / cells
  # home atoms anywhere; partners at distance cut*(1 + k ulp) in a random direction, so the
  # pair sits in the home cell's 3^d stencil; wrap partners into [0, L)
  a = (rng.random((n, dim)) * L)
  u = rng.normal(size=(n, dim))
  u /= np.linalg.norm(u, axis=1, keepdims=True)
  eps = np.finfo(dtype).eps
  scale = 1.0 + rng.integers(-8, 9, (n, 1)) * eps * rng.choice([0, 1, 4, 32, 1e3, 1e5], (n, 1))
  scale = np.where(rng.random((n, 1)) < 0.2, rng.random((n, 1)) * 1.2, scale)   # and bulk distances
  b = a + u * cut * scale
  a = a.astype(dtype)
  bw = np.mod(b, L).astype(dtype)
  ok = (bw >= 0).all(1) & (bw < dtype(L)).all(1) & (a < dtype(L)).all(1)        # regular atoms only
  a, bw = a[ok], bw[ok]
  # image shift of the stencil cell, as the kernel derives it from the cell coordinates
  ca = np.minimum((a / dtype(cs)).astype(np.int64), cells - 1)
  cb = np.minimum((bw / dtype(cs)).astype(np.int64), cells - 1)
  dc = cb - ca
  shift = np.where(dc > cells // 2, L, np.where(dc < -(cells // 2), -L, 0.0))   # b's cell is a wrapped image
  stencil = (np.abs(dc - np.sign(shift) * 0 - np.round(dc / cells) * cells) <= 1).all(1)
  a, bw, shift = a[stencil], bw[stencil], shift[stencil]
  assert len(a) > 100_000
  a2 = _filter_a2(a, bw, shift, dtype).astype(np.float64)
  d_fwd = _exact_d2(a, bw, L, dtype).astype(np.float64)
  d_rev = _exact_d2(bw, a, L, dtype).astype(np.float64)
  near = np.abs(d_fwd - cutoff_sq) < 50 * band                                  # where the decision is made
  worst = max(np.abs(a2 - d_fwd)[near].max(), np.abs(a2 - d_rev)[near].max())
  assert worst < 0.5 * band, (worst, band)
  # decisions: outside the band the filter's verdict equals the exact tests'
  lo, hi = cutoff_sq - band, cutoff_sq + band
  acc, rej = a2 < lo, a2 > hi
  assert (d_fwd[acc] < cutoff_sq).all() and (d_rev[acc] < cutoff_sq).all()
  assert (d_fwd[rej] >= cutoff_sq).all() and (d_rev[rej] >= cutoff_sq).all()


def _general_d2(sa, sb, side, dtype):
  """space.periodic_general with unit-cube positions (space.py:419-433) in separately rounded
  ops: box * (mod(fl(sa - sb) + 1/2, 1) - 1/2), sum of squares."""
  d = (sa - sb).astype(dtype)
  m = (np.mod((d + dtype(0.5)).astype(dtype), dtype(1.0)).astype(dtype) - dtype(0.5)).astype(dtype)
  g = (m * side.astype(dtype)).astype(dtype)
  sq = (g * g).astype(dtype)
  acc = sq[:, 0]
  for k in range(1, sq.shape[1]):
    acc = (acc + sq[:, k]).astype(dtype)
  return acc


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('L,cells', [(33.592, 11), (105.81, 37)])
def test_prefilter_band_for_unit_cube_positions(dtype, L, cells):
  """periodic_general: the metric is evaluated on unit-cube coordinates, the pre-filter on the
  real-space sorted copy fl(s * side) -- one more rounding of ulp(L)/2 per coordinate than the
  orthorhombic derivation counts.  fill() doubles the band for general spaces; the emulated error
  stays inside it."""
  rng = np.random.default_rng(int(L) + 1)
  dim, cut = 3, 2.8
  side = np.array([L, L * 1.07, L * 0.93])
  cutoff_sq = float(dtype(cut) ** 2)
  band = 2.0 * _band(float(side.max()), cutoff_sq, dim, dtype)      # P.band of a general space
  n = 400_000
  a_real = rng.random((n, dim)) * side
  u = rng.normal(size=(n, dim))
  u /= np.linalg.norm(u, axis=1, keepdims=True)
  eps = np.finfo(dtype).eps
  scale = 1.0 + rng.integers(-8, 9, (n, 1)) * eps * rng.choice([0, 1, 4, 32, 1e3, 1e5], (n, 1))
  scale = np.where(rng.random((n, 1)) < 0.2, rng.random((n, 1)) * 1.2, scale)
  b_real = a_real + u * cut * scale
  sa = (a_real / side).astype(dtype)
  sb = np.mod(b_real / side, 1.0).astype(dtype)
  ok = (sb >= 0).all(1) & (sb < 1).all(1) & (sa < 1).all(1)
  sa, sb = sa[ok], sb[ok]
  # the cell-sorted copy the pre-filter reads: real = fl(s * side)
  ra = (sa * side.astype(dtype)).astype(dtype)
  rb = (sb * side.astype(dtype)).astype(dtype)
  cs = 1.0 / cells                                                   # unit-cube grid
  ca = np.minimum((sa / dtype(cs)).astype(np.int64), cells - 1)
  cb = np.minimum((sb / dtype(cs)).astype(np.int64), cells - 1)
  dc = cb - ca
  wrap = np.where(dc > cells // 2, 1.0, np.where(dc < -(cells // 2), -1.0, 0.0))
  stencil = (np.abs(dc - np.round(dc / cells) * cells) <= 1).all(1)
  ra, rb, sa, sb, wrap = ra[stencil], rb[stencil], sa[stencil], sb[stencil], wrap[stencil]
  assert len(ra) > 100_000
  a2 = _filter_a2(ra, rb, wrap * side, dtype).astype(np.float64)    # home shifted by +-side (real space)
  d_fwd = _general_d2(sa, sb, side, dtype).astype(np.float64)
  d_rev = _general_d2(sb, sa, side, dtype).astype(np.float64)
  near = np.abs(d_fwd - cutoff_sq) < 50 * band
  worst = max(np.abs(a2 - d_fwd)[near].max(), np.abs(a2 - d_rev)[near].max())
  assert worst < 0.5 * band, (worst, band)
  lo, hi = cutoff_sq - band, cutoff_sq + band
  acc, rej = a2 < lo, a2 > hi
  assert (d_fwd[acc] < cutoff_sq).all() and (d_rev[acc] < cutoff_sq).all()
  assert (d_fwd[rej] >= cutoff_sq).all() and (d_rev[rej] >= cutoff_sq).all()
