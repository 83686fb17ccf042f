"""space.periodic_general with a full-matrix (triclinic) box: neighbour lists element-exact against
the oracle (whose matrix-vector products sum j = 0, 1, 2 like the kernels; XLA's own order is
unpinned), pair and Stillinger-Weber energies / forces, a short NVE trajectory, and the lattice
identity: a cubic cell sheared by one full box vector is the same crystal."""
import functools

import numpy as np
import pytest
import torch

from oracle import energy as oenergy
from oracle import partition as opart
from oracle import simulate as osim
from oracle import space as ospace
from tests import util

pytestmark = pytest.mark.gpu


def _jmd():
  import jax_md_b200 as jmd
  return jmd


def _dev(x):
  return torch.as_tensor(x, device='cuda')


def _tric(L, dtype=np.float32, dim=3):
  H = np.array([[L, 0.25 * L, -0.15 * L], [0, 0.95 * L, 0.2 * L], [0, 0, 1.05 * L]], dtype)
  return H[:dim, :dim].copy()


@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
@pytest.mark.parametrize('dim,dtype', [(3, np.float32), (3, np.float64), (2, np.float64)])
def test_triclinic_neighbor_lists_match_oracle(fmt, dim, dtype):
  jmd = _jmd()
  rng = np.random.default_rng(5)
  L = 12.0 if dim == 3 else 30.0
  H = _tric(L, dtype, dim)
  N = 1500 if dim == 3 else 800
  S = rng.random((N, dim)).astype(dtype)
  d_o, _ = ospace.periodic_general(H)
  d_g, _ = jmd.space.periodic_general(H)
  nf_o = opart.neighbor_list(d_o, H, np.float32(2.0), np.float32(0.3), fractional_coordinates=True,
                             format=opart.Format[fmt])
  nf_g = jmd.partition.neighbor_list(d_g, H, np.float32(2.0), np.float32(0.3), fractional_coordinates=True,
                                     format=jmd.partition.NeighborListFormat[fmt])
  nb_o, nb_g = nf_o.allocate(S), nf_g.allocate(_dev(S))
  assert nb_g._ws.c.space.triclinic == 1 and nb_g._ws.c.use_cells == 1
  np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)
  assert nb_g.max_occupancy == nb_o.max_occupancy and nb_g.cell_list_capacity == nb_o.cell_list_capacity
  assert int(nb_g.error.code) == int(nb_o.error)
  S2 = np.mod(S + rng.normal(0, 0.02, S.shape).astype(dtype), 1.0).astype(dtype)
  nb_o, nb_g = nb_o.update(S2), nb_g.update(_dev(S2))
  np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)
  np.testing.assert_array_equal(nb_g.reference_position.cpu().numpy(), nb_o.reference_position)


@pytest.mark.parametrize('fmt', ['Dense', 'OrderedSparse'])
def test_triclinic_real_space_positions_all_pairs(fmt):
  """fractional_coordinates=False with a matrix box: no cell grid (like the reference), all-pairs."""
  jmd = _jmd()
  rng = np.random.default_rng(6)
  H = _tric(9.0, np.float64)
  S = rng.random((300, 3))
  X = S @ H.T
  d_o, _ = ospace.periodic_general(H, fractional_coordinates=False)
  d_g, _ = jmd.space.periodic_general(H, fractional_coordinates=False)
  # (no disable_cell_list needed: partition.py:1052 never builds a grid for a matrix box)
  nf_o = opart.neighbor_list(d_o, H, np.float32(2.0), np.float32(0.3), format=opart.Format[fmt])
  nf_g = jmd.partition.neighbor_list(d_g, H, np.float32(2.0), np.float32(0.3),
                                     format=jmd.partition.NeighborListFormat[fmt])
  nb_o, nb_g = nf_o.allocate(X), nf_g.allocate(_dev(X))
  assert nb_g._ws.c.use_cells == 0 and nb_g.cell_list_capacity is None and nb_o.cell_list_capacity is None
  np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_triclinic_lj_energy_forces_and_nve(dtype):
  jmd = _jmd()
  R, L = util.fcc(6, dtype=np.float64)
  H = _tric(L, dtype)
  S = np.mod(util.jitter(R, L, 0.04) / L, 1.0).astype(dtype)      # the fcc sites, sheared with the box
  N = len(S)
  d_o, s_o = ospace.periodic_general(H)
  d_g, s_g = jmd.space.periodic_general(H)
  nf_o = opart.neighbor_list(d_o, H, np.float32(2.5), np.float32(0.3), fractional_coordinates=True,
                             format=opart.Dense, capacity_multiplier=1.5)
  nf_g, efn = jmd.energy.lennard_jones_neighbor_list(d_g, H, dr_threshold=0.3, fractional_coordinates=True,
                                                     format=jmd.partition.Dense, capacity_multiplier=1.5)
  pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
  nb_o = nf_o.allocate(S)
  Sd = _dev(S)
  nb_g = nf_g.allocate(Sd)
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(pot, d_o, S.astype(np.float64), nb_o, want_grads=True,
                                                  sigma=np.float64(1.0), epsilon=np.float64(1.0))
  rt = 2e-5 if dtype == np.float32 else 1e-10
  np.testing.assert_allclose(float(efn(Sd, neighbor=nb_g)), E_o, rtol=rt)
  Fg = jmd.quantity.force(efn)(Sd, neighbor=nb_g).cpu().numpy()
  np.testing.assert_allclose(Fg, F_o, rtol=rt, atol=rt * np.abs(F_o).max())
  # virial against the oracle's strain derivative
  W_o = oenergy.pair_virial(pot, d_o, S.astype(np.float64), nb_o, sigma=np.float64(1.0), epsilon=np.float64(1.0))
  np.testing.assert_allclose(efn.virial(Sd, neighbor=nb_g).cpu().numpy(), W_o, rtol=rt * 10,
                             atol=rt * 10 * np.abs(W_o).max())
  # NVE: unit-cube positions, real-space momenta; the shift goes through the inverse box
  P = util.momenta(N, 3, kT=0.5, dtype=dtype)
  hold = {'nb': nb_o}

  def f_o(X):
    return oenergy.pair_neighbor_list_energy(pot, d_o, X, hold['nb'], want_grads=True,
                                             sigma=dtype(1.0), epsilon=dtype(1.0))[1]
  init_o, step_o = osim.nve(f_o, s_o, 1e-3)
  st_o = init_o(S, P, mass=dtype(1.0))
  init_g, step_g = jmd.simulate.nve(efn, s_g, 1e-3)
  st_g = init_g(0, Sd, kT=0.5, momenta=_dev(P), neighbor=nb_g)
  for _ in range(60):
    hold['nb'] = hold['nb'].update(st_o.position)
    st_o = step_o(st_o)
    nb_g = nb_g.update(st_g.position)
    st_g = step_g(st_g, neighbor=nb_g)
  dX = st_g.position.cpu().numpy() - st_o.position
  dX -= np.round(dX)
  assert np.abs(dX).max() < (2e-5 if dtype == np.float32 else 1e-10)
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum,
                             atol=2e-3 if dtype == np.float32 else 1e-8, rtol=0)


def _shear_equivalent(R, L):
  """The cubic cell L*I and the sheared cell [[L, L, 0], [0, L, L], [0, 0, L]] generate the same
  lattice of images: unit-cube coordinates of the same atoms in the sheared cell."""
  H = np.array([[L, L, 0.0], [0.0, L, L], [0.0, 0.0, L]])
  S = np.mod(R @ np.linalg.inv(H).T, 1.0)
  return H, S


def test_sheared_cell_is_the_same_crystal_lj():
  jmd = _jmd()
  R, L = util.fcc(6, dtype=np.float64)
  Rj = np.mod(util.jitter(R, L, 0.05, seed=2), L)
  H, S = _shear_equivalent(Rj, L)
  d_c, _ = jmd.space.periodic(np.float64(L))
  nf_c, e_c = jmd.energy.lennard_jones_neighbor_list(d_c, np.float64(L), dr_threshold=0.3,
                                                     format=jmd.partition.Dense, capacity_multiplier=1.5)
  d_t, _ = jmd.space.periodic_general(H)
  nf_t, e_t = jmd.energy.lennard_jones_neighbor_list(d_t, H, dr_threshold=0.3, fractional_coordinates=True,
                                                     format=jmd.partition.Dense, capacity_multiplier=1.5)
  Rc, St = _dev(Rj), _dev(S)
  nb_c, nb_t = nf_c.allocate(Rc), nf_t.allocate(St)
  assert not bool(nb_t.did_buffer_overflow)
  np.testing.assert_allclose(float(e_t(St, neighbor=nb_t)), float(e_c(Rc, neighbor=nb_c)), rtol=1e-11)
  np.testing.assert_allclose(jmd.quantity.force(e_t)(St, neighbor=nb_t).cpu().numpy(),
                             jmd.quantity.force(e_c)(Rc, neighbor=nb_c).cpu().numpy(), rtol=1e-8, atol=1e-9)
  np.testing.assert_array_equal(np.sort(nb_t.idx.cpu().numpy(), -1), np.sort(nb_c.idx.cpu().numpy(), -1))


def test_sheared_cell_is_the_same_crystal_sw_and_nvt_runs():
  jmd = _jmd()
  R, L = util.diamond(4, a=5.431, dtype=np.float64)
  Rj = np.mod(util.jitter(R, L, 0.05, seed=3), L)
  H, S = _shear_equivalent(Rj, L)
  d_c, _ = jmd.space.periodic(np.float64(L))
  nf_c, e_c = jmd.energy.stillinger_weber_neighbor_list(d_c, np.float64(L))
  d_t, s_t = jmd.space.periodic_general(H)
  nf_t, e_t = jmd.energy.stillinger_weber_neighbor_list(d_t, H, fractional_coordinates=True)
  Rc, St = _dev(Rj), _dev(S)
  nb_c, nb_t = nf_c.allocate(Rc, extra_capacity=4), nf_t.allocate(St, extra_capacity=4)
  np.testing.assert_allclose(float(e_t(St, neighbor=nb_t)), float(e_c(Rc, neighbor=nb_c)), rtol=1e-11)
  np.testing.assert_allclose(jmd.quantity.force(e_t)(St, neighbor=nb_t).cpu().numpy(),
                             jmd.quantity.force(e_c)(Rc, neighbor=nb_c).cpu().numpy(), rtol=1e-8, atol=1e-9)
  np.testing.assert_allclose(e_t.virial(St, neighbor=nb_t).cpu().numpy(),
                             e_c.virial(Rc, neighbor=nb_c).cpu().numpy(), rtol=1e-8, atol=1e-8)
  unit = jmd.units.metal_unit_system()
  dt, kT = 1e-3 * unit['time'], 300.0 * unit['temperature']
  init, step = jmd.simulate.nvt_nose_hoover(e_t, s_t, dt, kT, chain_length=3, chain_steps=1, sy_steps=1, tau=100 * dt)
  st = init(0, St, mass=28.0855, neighbor=nb_t)
  _, s_c = jmd.space.periodic(np.float64(L))
  init_c, step_c = jmd.simulate.nvt_nose_hoover(e_c, s_c, dt, kT, chain_length=3, chain_steps=1, sy_steps=1,
                                                tau=100 * dt)
  st_c = init_c(0, Rc, mass=28.0855, neighbor=nb_c)             # same key: same real-space momenta
  inv0 = float(jmd.simulate.nvt_nose_hoover_invariant(e_t, st, kT, neighbor=nb_t))
  for _ in range(100):
    nb_t = nb_t.update(st.position)
    st = step(st, neighbor=nb_t)
    nb_c = nb_c.update(st_c.position)
    st_c = step_c(st_c, neighbor=nb_c)
  # the same trajectory in both descriptions of the cell
  np.testing.assert_allclose(st.momentum.cpu().numpy(), st_c.momentum.cpu().numpy(), atol=1e-9, rtol=0)
  dR = st.position.cpu().numpy() @ H.T - st_c.position.cpu().numpy()
  assert np.abs(dR - np.round(dR / L) * L).max() < 1e-9
  assert not bool(nb_t.did_buffer_overflow)
  assert bool(((st.position >= 0) & (st.position < 1)).all())
  inv1 = float(jmd.simulate.nvt_nose_hoover_invariant(e_t, st, kT, neighbor=nb_t))
  assert abs(inv1 - inv0) / len(R) < 2e-4, (inv0, inv1)      # velocity Verlet at dt = 1 fs: O(dt^2) wobble
