"""Pins the CPU oracle against the reference's own golden vectors (CPU only).

Sources of every number: tests/golden/make_golden.py.
"""
import json
import os

import numpy as np
import pytest

from oracle import energy as oenergy
from oracle import partition as opart
from oracle import simulate as osim
from oracle import space as ospace
from tests import util

with open(os.path.join(util.GOLDEN, 'goldens.json')) as f:
  G = json.load(f)


# -- cell list goldens (reference tests/partition_test.py:61-82) ---------------

def test_cell_list_emplace_2d_golden():
  for dtype in (np.float32, np.float64):
    box = np.array([8.65, 8.0], np.float32)
    R = np.array([[0.25, 0.25], [8.5, 1.95], [8.1, 1.5], [3.7, 7.9]], dtype)
    cl = opart.cell_list_build(R, box, np.float32(1.0))
    cps = cl.cells_per_side
    assert list(cps) == [8, 8]
    ids = cl.id_buffer.reshape(cps[1], cps[0], -1)        # [cy, cx, cap]
    pos = cl.position_buffer.reshape(cps[1], cps[0], -1, 2)
    # NOTE the reference asserts [1, 8, 1]: with box 8.65 / cell 1.0 there are
    # 8 cells per side, so x index 8 == the reference test's own indexing of
    # a [8, 8, cap] buffer is out of range in NumPy; JAX clamps it to 7.
    assert ids[0, 0, 0] == 0
    assert ids[1, 7, 1] == 1
    assert ids[1, 7, 0] == 2
    assert ids[7, 3, 1] == 3
    np.testing.assert_allclose(pos[1, 7, 1], R[1])
    np.testing.assert_allclose(pos[7, 3, 1], R[3])
    flat = cl.id_buffer.reshape(-1)
    out = np.zeros((5, 2), dtype)
    out[flat] = cl.position_buffer.reshape(-1, 2)
    np.testing.assert_allclose(out[:-1], R)


@pytest.mark.parametrize('dim', [2, 3])
def test_cell_list_random_emplace(dim):
  rng = np.random.default_rng(1)
  R = (9.0 * rng.random((1000, dim))).astype(np.float32)
  cl = opart.cell_list_build(R, np.float32(9.0), np.float32(1.0))
  flat = cl.id_buffer.reshape(-1)
  out = np.zeros((1001, dim), np.float32)
  out[flat] = cl.position_buffer.reshape(-1, dim)
  np.testing.assert_array_equal(out[:-1], R)
  assert not cl.did_buffer_overflow


# -- capacity goldens (reference tests/partition_test.py:488-546) ---------------

@pytest.mark.parametrize('case', [(0.12, True, 1.5), (0.25, False, 1.5),
                                  (0.31, False, 1.5), (0.31, False, 1.0)])
@pytest.mark.parametrize('mask_self', [False, True])
@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
def test_issue191_shapes(case, mask_self, fmt):
  r_cut, disable, cm = case
  box = np.ones(3)
  R = np.ones((20, 3)) * 0.5
  if fmt == 'Dense':
    want = (20, 19) if mask_self else (20, 20)
  elif fmt == 'Sparse':
    want = (2, 380) if mask_self else (2, 400)
  else:
    want = (2, 190)
  d, _ = ospace.periodic(box)
  nf = opart.neighbor_list(d, box, r_cut, 0.1 * r_cut, capacity_multiplier=cm,
                           disable_cell_list=disable, mask_self=mask_self,
                           format=opart.Format[fmt])
  nbrs = nf.allocate(R)
  assert not nbrs.did_buffer_overflow
  assert nbrs.idx.shape == want
  new = nbrs.update(R + 0.1)
  assert not new.did_buffer_overflow
  assert new.idx.shape == want


def test_cell_list_overflow_flag():
  """reference tests/partition_test.py:363-401."""
  d, _ = ospace.free()
  nf = opart.neighbor_list(d, 100.0, 3.0, 0.0)
  R = np.array([[20., 20.], [30., 30.], [40., 40.], [50., 50.]], np.float32)
  nbrs = nf.allocate(R)
  assert nbrs.idx.dtype == np.int32
  R2 = np.array([[20., 20.], [20., 20.], [40., 40.], [50., 50.]], np.float32)
  nbrs = nbrs.update(R2)
  assert nbrs.did_buffer_overflow


# -- neighbour list == brute force (reference tests/partition_test.py:203-301) --

@pytest.mark.parametrize('dim', [2, 3])
@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
def test_neighbor_list_build_matches_bruteforce(dim, fmt):
  rng = np.random.default_rng(0)
  box = np.array([9.0, 4.0, 7.25][:dim], np.float32)
  N = 600
  R = (rng.random((N, dim)) * box).astype(np.float32)
  d, _ = ospace.periodic(box)
  cutoff = 1.23
  nf = opart.neighbor_list(d, box, cutoff, 0.0, capacity_multiplier=1.1,
                           format=opart.Format[fmt])
  nbrs = nf.allocate(R)
  assert nbrs.use_cell_list
  dR = d(R[:, None, :], R[None, :, :])
  d2 = ospace.square_distance(dR)
  want = (d2 < np.float32(cutoff ** 2)) & ~np.eye(N, dtype=bool)
  if fmt == 'Dense':
    got = np.zeros((N, N), bool)
    rows = np.broadcast_to(np.arange(N)[:, None], nbrs.idx.shape)
    m = nbrs.idx < N
    got[rows[m], nbrs.idx[m]] = True
    # Dense keeps a pair only if both orientations pass (two-stage test)
    np.testing.assert_array_equal(got, want & want.T)
  else:
    pairs = util.sparse_pairs(nbrs.idx, N)
    got = np.zeros((N, N), bool)
    got[pairs[:, 0], pairs[:, 1]] = True
    if fmt == 'OrderedSparse':
      want = want & (np.arange(N)[None, :] < np.arange(N)[:, None])
    np.testing.assert_array_equal(got, want)


# -- energy goldens -------------------------------------------------------------

@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
def test_jammed_soft_sphere_energy_golden(fmt):
  s = np.load(os.path.join(util.GOLDEN, 'jammed_state.npz'))
  R = s['real_position']
  L = s['box'][0, 0]
  d, _ = ospace.periodic(L)
  pot = oenergy.PairPotential('soft_sphere')
  E_bf = oenergy.pair_energy_bruteforce(pot, d, R, species=s['species'],
                                        sigma=s['sigma'], epsilon=1.0, alpha=2.0)
  np.testing.assert_allclose(E_bf, G['jammed_energy'], rtol=1e-12)
  nf = opart.neighbor_list(d, L, np.max(s['sigma']), 0.2,
                           format=opart.Format[fmt])
  nbrs = nf.allocate(R)
  E = oenergy.pair_neighbor_list_energy(pot, d, R, nbrs, species=s['species'],
                                        sigma=s['sigma'], epsilon=1.0, alpha=2.0)
  np.testing.assert_allclose(E, G['jammed_energy'], rtol=1e-12)


def test_lammps_lj_energy_golden():
  s = np.load(os.path.join(util.GOLDEN, 'lammps_lj.npz'))
  box = np.float32(s['box'])
  R = s['R'] * box                       # xs ys zs are fractional
  d, _ = ospace.periodic(box)
  dr = ospace.distance(d(R[:, None, :], R[None, :, :]))
  U = np.where(dr < np.float32(2.5), oenergy.lennard_jones(dr), 0.0)
  np.fill_diagonal(U, 0.0)
  E = U.sum() / 2 / len(R)
  np.testing.assert_allclose(E, G['lammps_lj_energy_per_atom'], rtol=5e-5,
                             atol=5e-5)


@pytest.mark.parametrize('n', [2, 3])
def test_stillinger_weber_golden(n):
  R, L = util.diamond(n, a=G['sw_lattice_constant'])
  d, _ = ospace.periodic(L)
  if n == 2:     # box too small for cells: brute-force candidates
    nf = opart.neighbor_list(d, L, 3.77118, 0.5, format=opart.Dense)
  else:
    nf = opart.neighbor_list(d, L, 3.77118, 0.5, format=opart.Dense)
  nbrs = nf.allocate(R)
  E = oenergy.stillinger_weber_energy(d, R, nbrs)
  np.testing.assert_allclose(E / len(R), G['sw_diamond_energy_per_atom'],
                             rtol=1e-12)


# -- oracle closed-form forces vs finite differences of the oracle energy --------

def _fd_force(efn, R, h=1e-6):
  F = np.zeros_like(R)
  for i in range(R.shape[0]):
    for k in range(R.shape[1]):
      Rp, Rm = R.copy(), R.copy()
      Rp[i, k] += h
      Rm[i, k] -= h
      F[i, k] = -(efn(Rp) - efn(Rm)) / (2 * h)
  return F


@pytest.mark.parametrize('kind', ['lj', 'soft_sphere', 'morse'])
@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
def test_pair_forces_match_finite_differences(kind, fmt):
  R, L = util.fcc(3, rho=0.8, dtype=np.float64)
  R = util.jitter(R, L, 0.05)
  L = float(L)
  d, _ = ospace.periodic(L)
  if kind == 'lj':
    pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
    params = dict(sigma=np.float64(1.0), epsilon=np.float64(1.0))
    rc = 2.5
  elif kind == 'morse':
    pot = oenergy.PairPotential('morse', np.float32(2.0), np.float32(2.5))
    params = dict(sigma=np.float64(1.0), epsilon=np.float64(5.0), alpha=np.float64(5.0))
    rc = 2.5
  else:
    pot = oenergy.PairPotential('soft_sphere')
    params = dict(sigma=np.float64(1.3), epsilon=np.float64(1.0), alpha=np.float64(2.0))
    rc = 1.3
  nf = opart.neighbor_list(d, L, rc, 0.3, format=opart.Format[fmt])
  nbrs = nf.allocate(R)
  E, F, dp = oenergy.pair_neighbor_list_energy(pot, d, R, nbrs, want_grads=True,
                                               **params)
  efn = lambda Rx: oenergy.pair_neighbor_list_energy(pot, d, Rx, nbrs, **params)
  Ffd = _fd_force(efn, R)[:6]
  np.testing.assert_allclose(F[:6], Ffd, rtol=2e-6, atol=2e-6)
  for name in ('sigma', 'epsilon'):
    h = 1e-6
    pp, pm = dict(params), dict(params)
    pp[name] += h
    pm[name] -= h
    fd = (oenergy.pair_neighbor_list_energy(pot, d, R, nbrs, **pp) -
          oenergy.pair_neighbor_list_energy(pot, d, R, nbrs, **pm)) / (2 * h)
    np.testing.assert_allclose(dp[name], fd, rtol=1e-5, atol=1e-6)


def test_sw_forces_match_finite_differences():
  R, L = util.diamond(2)
  R = util.jitter(R, L, 0.08, seed=3)
  d, _ = ospace.periodic(L)
  nf = opart.neighbor_list(d, L, 3.77118, 0.5, format=opart.Dense)
  nbrs = nf.allocate(R)
  E, F = oenergy.stillinger_weber_energy(d, R, nbrs, want_force=True)
  efn = lambda Rx: oenergy.stillinger_weber_energy(d, Rx, nbrs)
  Ffd = _fd_force(efn, R)
  np.testing.assert_allclose(F, Ffd, rtol=2e-6, atol=2e-6)


# -- integrators: conservation on the oracle itself ------------------------------

def _lj_system(n=3, dtype=np.float64):
  R, L = util.fcc(n, dtype=dtype)
  L = dtype(L)
  d, s = ospace.periodic(L)
  pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
  nf = opart.neighbor_list(d, L, 2.5, 0.0, disable_cell_list=True,
                           format=opart.Dense)

  def both(Rx):
    nbrs = nf.allocate(Rx)
    return oenergy.pair_neighbor_list_energy(pot, d, Rx, nbrs, want_grads=True,
                                             sigma=1.0, epsilon=1.0)
  return R, d, s, both


def test_nve_conserves_energy():
  R, d, s, both = _lj_system()
  P = util.momenta(len(R), 3, kT=0.5, dtype=np.float64)
  init, step = osim.nve(lambda Rx: both(Rx)[1], s, 1e-3)
  st = init(R, P)
  E0 = both(st.position)[0] + osim.kinetic_energy(st.momentum, st.mass)
  for _ in range(50):
    st = step(st)
  E1 = both(st.position)[0] + osim.kinetic_energy(st.momentum, st.mass)
  assert abs(E1 - E0) < 1e-5 * abs(E0)


@pytest.mark.parametrize('sy', [1, 3, 5, 7])
def test_nvt_invariant(sy):
  R, d, s, both = _lj_system()
  kT = 0.7
  P = util.momenta(len(R), 3, kT=kT, dtype=np.float64)
  init, step = osim.nvt_nose_hoover(lambda Rx: both(Rx)[1], s, 1e-3, kT,
                                    sy_steps=sy)
  st = init(R, P)
  H0 = osim.nvt_nose_hoover_invariant(both(st.position)[0], st, kT)
  for _ in range(40):
    st = step(st)
  H1 = osim.nvt_nose_hoover_invariant(both(st.position)[0], st, kT)
  assert abs(H1 - H0) < 1e-5 * abs(H0)


def test_fire_descent_reduces_force():
  R, d, s, both = _lj_system()
  R = util.jitter(R, 1e9, 0.05)
  init, step = osim.fire_descent(lambda Rx: both(Rx)[1], s, dt_start=0.01,
                                 dt_max=0.04)
  st = init(R)
  f0 = np.abs(st.force).max()
  for _ in range(150):
    st = step(st)
  assert np.abs(st.force).max() < 0.05 * f0


def test_pressure_and_stress_goldens():
  """Reference tests/quantity_test.py:134-150 (jammed soft spheres, P =
  0.06307342050945483) and :436-455 (LAMMPS LJ stress incl. the kinetic term)."""
  s = np.load(os.path.join(util.GOLDEN, 'jammed_state.npz'))
  R = s['real_position']
  L = float(s['box'][0, 0])
  d, _ = ospace.periodic(L)
  nb = opart.neighbor_list(d, L, np.float64(np.max(s['sigma'])), np.float64(0.0),
                           format=opart.Dense).allocate(R)
  pot = oenergy.PairPotential('soft_sphere')
  P = oenergy.pressure(pot, d, R, L, nb, species=s['species'], sigma=s['sigma'],
                       epsilon=np.float64(1.0), alpha=np.float64(2.0))
  np.testing.assert_allclose(P, G['jammed_pressure'], rtol=1e-9)

  s = np.load(os.path.join(util.GOLDEN, 'lammps_lj.npz'))
  box = np.float32(s['box'])
  R = (s['R'] * box).astype(np.float64)
  r = s['stress_row']
  C = np.array([[r[0], r[3], r[4]], [r[3], r[1], r[5]], [r[4], r[5], r[2]]])
  d, _ = ospace.periodic(box)
  for fmt in (opart.Dense, opart.Sparse, opart.OrderedSparse):
    nb = opart.neighbor_list(d, box, np.float32(2.5), np.float32(0.0), format=fmt).allocate(R)
    S = oenergy.stress(oenergy.PairPotential('lj'), d, R, box, nb, velocity=s['V'],
                       sigma=np.float64(1.0), epsilon=np.float64(1.0))
    np.testing.assert_allclose(S, C, rtol=5e-5, atol=5e-5)
