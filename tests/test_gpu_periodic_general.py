"""space.periodic_general + fractional coordinates + `box=` (reference
space.py:332-472, partition.py:595-638, 1045-1051, 1125-1139) against the oracle."""
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


def _boxes(L):
  return {'scalar': np.float32(L), 'vector': np.array([L, L * 1.1, L * 0.95], np.float32),
          'matrix': np.diag(np.array([L, L * 1.05, L * 0.9], np.float32))}


def _same(nb_g, nb_o, fmt, N):
  if fmt == 'Dense':
    assert nb_g.idx.shape == nb_o.idx.shape
    np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)        # element-exact, order included
  else:
    np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)
  assert nb_g.max_occupancy == nb_o.max_occupancy
  assert nb_g.cell_list_capacity == nb_o.cell_list_capacity
  assert int(nb_g.error.code) == int(nb_o.error)


@pytest.mark.parametrize('kind', ['scalar', 'vector', 'matrix'])
@pytest.mark.parametrize('fmt', ['Dense', 'Sparse', 'OrderedSparse'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_fractional_neighbor_lists_match_oracle(kind, fmt, dtype):
  jmd = _jmd()
  R, L = util.fcc(7, dtype=np.float64)
  box = _boxes(L)[kind]
  diag = np.diag(box) if box.ndim == 2 else np.broadcast_to(box, (3,))
  S = np.mod(util.jitter(R, L, 0.06) / L, 1.0).astype(dtype)              # unit-cube positions
  N = len(S)
  d_o, _ = ospace.periodic_general(box)
  d_g, _ = jmd.space.periodic_general(box)
  F = jmd.partition.NeighborListFormat[fmt]
  nf_o = opart.neighbor_list(d_o, box, np.float32(2.5), np.float32(0.3), fractional_coordinates=True,
                             format=opart.Format[fmt])
  nf_g = jmd.partition.neighbor_list(d_g, box, np.float32(2.5), np.float32(0.3), fractional_coordinates=True,
                                     format=F)
  nb_o = nf_o.allocate(S)
  nb_g = nf_g.allocate(_dev(S))
  _same(nb_g, nb_o, fmt, N)
  assert int(nb_g.error.code) & 8              # the MALFORMED_BOX quirk: set for a VALID box
  assert not bool(nb_g.did_buffer_overflow)
  # move atoms past the skin: both rebuild, lists agree again
  rng = np.random.default_rng(1)
  S2 = np.mod(S + (rng.normal(0, 0.2, S.shape) / diag).astype(dtype), 1.0).astype(dtype)
  nb_o = nb_o.update(S2)
  nb_g = nb_g.update(_dev(S2))
  _same(nb_g, nb_o, fmt, N)


def test_box_kwarg_error_bits_and_metric():
  """update(position, box=...): CELL_SIZE_TOO_SMALL when the new box needs bigger cells,
  the MALFORMED_BOX quirk, and the new box drives the metric of the rebuild."""
  jmd = _jmd()
  R, L = util.fcc(7, dtype=np.float64)
  S = np.mod(util.jitter(R, L, 0.05) / L, 1.0).astype(np.float32)
  box = np.float32(L)
  d_o, _ = ospace.periodic_general(box)
  d_g, _ = jmd.space.periodic_general(box)
  nf_o = opart.neighbor_list(d_o, box, np.float32(2.5), np.float32(0.3), fractional_coordinates=True,
                             format=opart.Dense)
  nf_g = jmd.partition.neighbor_list(d_g, box, np.float32(2.5), np.float32(0.3), fractional_coordinates=True)
  nb_o, nb_g = nf_o.allocate(S), nf_g.allocate(_dev(S))
  for scale in (1.04, 0.8):                      # a larger box (fine), then one that is too small
    b2 = np.float32(L * scale)
    # the skin predicate is evaluated with the NEW box: force a rebuild by moving one atom
    S2 = S.copy()
    S2[0, 0] = np.mod(S2[0, 0] + 0.3 / L, 1.0)
    nb_o = nb_o.update(S2, box=b2)
    nb_g = nb_g.update(_dev(S2), box=b2)
    assert int(nb_g.error.code) == int(nb_o.error)
    if not (int(nb_o.error) & 3):
      np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)
  assert int(nb_g.error.code) & 4               # CELL_SIZE_TOO_SMALL after the shrink
  # partition.py:1125-1130: update() with box= but without fractional coordinates raises (allocate does not)
  d_p, _ = jmd.space.periodic(box)
  nb_p = jmd.partition.neighbor_list(d_p, box, 2.5, 0.3).allocate(_dev(S * L), box=box)
  with pytest.raises(ValueError):
    nb_p.update(_dev(S * L), box=box)


@pytest.mark.parametrize('frac', [True, False])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_lj_energy_force_and_nve_general_box(frac, dtype):
  """LJ over periodic_general (vector box; unit-cube and real-space parametrisations):
  energy, forces (real-space forces also for fractional positions, space.py:171-186) and a
  100-step NVE trajectory vs the oracle."""
  jmd = _jmd()
  R, L = util.fcc(6, dtype=np.float64)
  box = np.array([L, L * 1.08, L * 0.96], np.float32).astype(dtype)
  Rr = np.mod(util.jitter(R, L, 0.05) * (box / L), box)
  X = (Rr / box if frac else Rr).astype(dtype)
  N = len(X)
  d_o, s_o = ospace.periodic_general(box, fractional_coordinates=frac)
  d_g, s_g = jmd.space.periodic_general(box, fractional_coordinates=frac)
  nf_o = opart.neighbor_list(d_o, box, np.float32(2.5), np.float32(0.3), fractional_coordinates=frac,
                             format=opart.Dense)
  nf_g, efn = jmd.energy.lennard_jones_neighbor_list(d_g, box, dr_threshold=0.3, fractional_coordinates=frac,
                                                     format=jmd.partition.Dense)
  pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
  nb_o = nf_o.allocate(X)
  Xd = _dev(X)
  nb_g = nf_g.allocate(Xd)
  np.testing.assert_array_equal(np.sort(nb_g.idx.cpu().numpy(), -1), np.sort(nb_o.idx, -1))
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(pot, d_o, X.astype(np.float64), nb_o, want_grads=True,
                                                  sigma=np.float64(1.0), epsilon=np.float64(1.0))
  # (the oracle's analytic gradient is w.r.t. the REAL displacement -- what the reference's custom
  #  JVP of space.transform, space.py:171-186, reports for unit-cube positions as well)
  rt = 1e-5 if dtype == np.float32 else 1e-10
  np.testing.assert_allclose(float(efn(Xd, neighbor=nb_g)), E_o, rtol=rt, atol=rt)
  Fg = jmd.quantity.force(efn)(Xd, neighbor=nb_g).cpu().numpy()
  np.testing.assert_allclose(Fg, F_o, rtol=rt, atol=rt * np.abs(F_o).max())
  # NVE: positions stay in their parametrisation, the shift takes real-space displacements
  P = util.momenta(N, 3, kT=1.0, dtype=dtype)
  holder = {'nb': nb_o}

  def f_o(Xx):
    F = oenergy.pair_neighbor_list_energy(pot, d_o, Xx, holder['nb'], want_grads=True,
                                          sigma=dtype(1.0), epsilon=dtype(1.0))[1]
    return F
  init_o, step_o = osim.nve(f_o, s_o, 1e-3)
  st_o = init_o(X, P, mass=dtype(1.0))
  init_g, step_g = jmd.simulate.nve(efn, s_g, 1e-3)
  st_g = init_g(0, Xd, kT=1.0, momenta=_dev(P), neighbor=nb_g)
  for _ in range(100):
    holder['nb'] = holder['nb'].update(st_o.position)
    st_o = step_o(st_o)
    nb_g = nb_g.update(st_g.position)
    st_g = step_g(st_g, neighbor=nb_g)
  period = 1.0 if frac else box
  dX = st_g.position.cpu().numpy() - st_o.position
  dX -= np.round(dX / period) * period
  assert np.abs(dX).max() < (2e-4 if dtype == np.float32 else 1e-9)
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum,
                             atol=2e-3 if dtype == np.float32 else 1e-8, rtol=0)


def test_sw_nvt_with_periodic_general_as_in_the_example():
  """examples/units/nvt_si_sw.py: space.periodic_general(latvec) with a diagonal 3x3 box,
  unit-cube positions, `box=` passed to allocate; SW energy equals the plain periodic run."""
  jmd = _jmd()
  R, L = util.diamond(4, a=5.431, dtype=np.float64)
  Rj = util.jitter(R, L, 0.05, seed=3)
  latvec = np.diag(np.array([L, L, L], np.float64))
  d_g, s_g = jmd.space.periodic_general(latvec)
  nf, efn = jmd.energy.stillinger_weber_neighbor_list(d_g, latvec, disable_cell_list=True)
  S = _dev((Rj / L).astype(np.float64))
  nb = nf.allocate(S, box=latvec, extra_capacity=2)
  d_p, _ = jmd.space.periodic(np.float64(L))
  nf_p, efn_p = jmd.energy.stillinger_weber_neighbor_list(d_p, np.float64(L))
  Rp = _dev(Rj)
  nb_p = nf_p.allocate(Rp, extra_capacity=2)
  np.testing.assert_allclose(float(efn(S, neighbor=nb)), float(efn_p(Rp, neighbor=nb_p)), rtol=1e-10)
  np.testing.assert_allclose(jmd.quantity.force(efn)(S, neighbor=nb).cpu().numpy(),
                             jmd.quantity.force(efn_p)(Rp, neighbor=nb_p).cpu().numpy(), rtol=1e-8, atol=1e-9)
  unit = jmd.units.metal_unit_system()
  dt, kT = 1e-3 * unit['time'], 300.0 * unit['temperature']
  init, step = jmd.simulate.nvt_nose_hoover(efn, s_g, dt, kT, chain_length=3, chain_steps=1, sy_steps=1, tau=100 * dt)
  st = init(0, S, mass=28.0855, neighbor=nb, kT=kT)
  for _ in range(50):
    st = step(st, neighbor=nb, kT=kT)
    nb = nb.update(st.position)
  assert not bool(nb.did_buffer_overflow)
  assert bool(((st.position >= 0) & (st.position < 1)).all())        # still unit-cube coordinates
  T = float(jmd.simulate.temperature(st)) / unit['temperature']
  assert 50 < T < 600


@pytest.mark.parametrize('fmt', ['Dense', 'OrderedSparse'])
def test_unit_cube_positions_on_a_real_space_grid(fmt):
  """tests/simulate_test.py:591-620 builds its list WITHOUT fractional_coordinates=True from a
  periodic_general displacement and unit-cube positions: the cell grid is sized for the 21.7 A box
  and every atom falls into its corner cell.  Same lists as the oracle, periodic images included."""
  jmd = _jmd()
  R, L = util.diamond(4, a=5.431, dtype=np.float64)
  S = np.mod(util.jitter(R, L, 0.05, seed=4) / L, 1.0)
  box = np.eye(3) * L
  d_o, _ = ospace.periodic_general(box)
  d_g, _ = jmd.space.periodic_general(box)
  nf_o = opart.neighbor_list(d_o, box, np.float32(3.77118), np.float32(0.5), format=opart.Format[fmt])
  nf_g = jmd.partition.neighbor_list(d_g, box, np.float32(3.77118), np.float32(0.5),
                                     format=jmd.partition.NeighborListFormat[fmt])
  nb_o, nb_g = nf_o.allocate(S, box=box), nf_g.allocate(_dev(S), box=box)
  # a MATRIX box: `all(cell_size < box / 3)` (partition.py:1052) is false for its zero elements
  assert nb_g._ws.c.use_cells == 0 and nb_o.cell_list_capacity is None and nb_g.cell_list_capacity is None
  np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)
  assert nb_g.max_occupancy == nb_o.max_occupancy
  # the same with a VECTOR box does build the (degenerate) grid: unit-cube positions, 21.7 A cells
  bv = np.full(3, L)
  d_o2, _ = ospace.periodic_general(bv)
  d_g2, _ = jmd.space.periodic_general(bv)
  nf_o2 = opart.neighbor_list(d_o2, bv, np.float32(3.77118), np.float32(0.5), format=opart.Format[fmt])
  nf_g2 = jmd.partition.neighbor_list(d_g2, bv, np.float32(3.77118), np.float32(0.5),
                                      format=jmd.partition.NeighborListFormat[fmt])
  nb_o2, nb_g2 = nf_o2.allocate(S), nf_g2.allocate(_dev(S))
  assert nb_g2._ws.c.use_cells == 1 and nb_g2.cell_list_capacity == nb_o2.cell_list_capacity
  np.testing.assert_array_equal(nb_g2.idx.cpu().numpy(), nb_o2.idx)
  counts = (nb_g.idx.cpu().numpy() < len(S)).sum(-1) if fmt == 'Dense' else None
  if counts is not None:
    assert counts.min() >= 4                 # every atom sees its 4 bonded neighbours, faces included
