"""GPU parity on the BASELINE.json configurations that round 1 left untested
(VERDICT r01 "what's weak" 1):

  config 1   examples/nve_neighbor_list.py: 2-D, f64, OrderedSparse, NVE -- trajectory
             vs the oracle; 2-D FIRE vs the oracle
  config 2   LJ fcc N=32,000, NVE 10^4 steps, Dense and Sparse: |dE|/N stated and
             compared with the C port's drift over the same run
  config 4   Stillinger-Weber inside nvt_nose_hoover (chain 3, chain_steps 1, sy 1,
             metal units) vs the oracle
  N=1M       neighbour sets (Dense + Sparse) vs the C port, bit-exact
"""
import numpy as np
import pytest
import torch

from oracle import cport
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


# ---- config 4: SW inside NVT ----------------------------------------------------------

@pytest.mark.parametrize('dtype,n', [(np.float32, 3), (np.float64, 4)])
def test_sw_inside_nvt_nose_hoover_matches_oracle(dtype, n):
  """examples/units/nvt_si_sw.py semantics: metal units, kT = 300 K, dt = 1 fs,
  tau = 100 dt, chain_length 3, chain_steps 1, sy_steps 1, Si mass; the fused SW
  kernel's half kick + KE feed the Nose-Hoover chain (simulate._Stepper 'sw' branch)."""
  jmd = _jmd()
  unit = jmd.units.metal_unit_system()
  dt = 1e-3 * unit['time']
  kT = 300.0 * unit['temperature']
  mass = 28.0855 * unit['mass']
  R, L = util.diamond(n, a=5.431, dtype=dtype)
  R = util.jitter(R, dtype(L), 0.05, seed=3)
  N = len(R)
  P = (util.momenta(N, 3, kT=1.0, seed=4, dtype=np.float64) * np.sqrt(mass * kT)).astype(dtype)
  Lf = np.float32(L)
  d_o, s_o = ospace.periodic(Lf)
  nf_o = opart.neighbor_list(d_o, Lf, 3.77118, 0.5, format=opart.Dense)
  holder = {'nb': nf_o.allocate(R)}

  def f_o(Rx):
    holder['nb'] = holder['nb'].update(Rx)
    return oenergy.stillinger_weber_energy(d_o, Rx, holder['nb'], want_force=True)[1].astype(Rx.dtype)
  init_o, step_o = osim.nvt_nose_hoover(f_o, s_o, dt, kT, chain_length=3, chain_steps=1,
                                        sy_steps=1, tau=100 * dt)
  st_o = init_o(R, P, mass=dtype(mass))
  d_g, s_g = jmd.space.periodic(Lf)
  nf_g, efn = jmd.energy.stillinger_weber_neighbor_list(d_g, Lf)
  init_g, step_g = jmd.simulate.nvt_nose_hoover(efn, s_g, dt, kT, chain_length=3, chain_steps=1,
                                                sy_steps=1, tau=100 * dt)
  Rd = _dev(R)
  nbrs = nf_g.allocate(Rd)
  st_g = init_g(0, Rd, mass=mass, momenta=_dev(P), neighbor=nbrs)
  np.testing.assert_allclose(st_g.force.cpu().numpy(), st_o.force,
                             rtol=1e-5 if dtype == np.float32 else 1e-10,
                             atol=(1e-5 if dtype == np.float32 else 1e-10) * np.abs(st_o.force).max())
  steps = 100
  for _ in range(steps):
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
  assert not bool(nbrs.did_buffer_overflow)
  rt = 2e-3 if dtype == np.float32 else 1e-7
  dR = st_g.position.cpu().numpy() - st_o.position
  dR -= np.round(dR / L) * L
  assert np.abs(dR).max() < (2e-4 if dtype == np.float32 else 1e-9)
  pscale = np.abs(st_o.momentum).max()
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum, atol=rt * pscale, rtol=0)
  np.testing.assert_allclose(st_g.chain.momentum.cpu().numpy(), st_o.chain.momentum,
                             rtol=rt, atol=rt * np.abs(st_o.chain.momentum).max())
  np.testing.assert_allclose(st_g.chain.position.cpu().numpy(), st_o.chain.position,
                             rtol=rt, atol=rt * np.abs(st_o.chain.position).max())
  np.testing.assert_allclose(float(st_g.chain.kinetic_energy), float(st_o.chain.kinetic_energy), rtol=rt)


# ---- config 2: N = 32,000, 10^4 steps, Dense and Sparse -------------------------------------

_CPORT_DRIFT = {}


def _cport_drift(R, L, P, dt, steps):
  """(E_end - E_0) / N of the C port over the same run (cached: both formats compare
  against one CPU run)."""
  if 'v' not in _CPORT_DRIFT:
    sysc = cport.LJSystem(R, L, r_cutoff=2.5, skin=0.3, r_onset=2.0, row_capacity=160)
    sysc.run(P, dt, 0)
    e0, _ = sysc.force()
    E0 = e0 + 0.5 * float((sysc.P.astype(np.float64) ** 2).sum())
    rebuilds = sysc.run(None, dt, steps)
    e1, _ = sysc.force()
    E1 = e1 + 0.5 * float((sysc.P.astype(np.float64) ** 2).sum())
    assert not sysc.overflow
    sysc.close()
    _CPORT_DRIFT['v'] = ((E1 - E0) / len(R), rebuilds)
  return _CPORT_DRIFT['v']


@pytest.mark.parametrize('fmt', ['Dense', 'Sparse'])
def test_lj_32k_nve_1e4_steps_drift_matches_c_port(fmt):
  """BASELINE config 2: LJ fcc N=32,000, rho=0.8442, rc=2.5, skin 0.3, T*=1.0, NVE 10^4
  steps through lax.fori_loop (CUDA-graph replay; the rebuild decision stays on the
  device).  Stated bound: |E_end - E_0| / N < 2e-3 (f32, dt=0.005), and the drift agrees
  with the C port of the reference path over the same run within the same bound."""
  jmd = _jmd()
  R, L = util.fcc(20, dtype=np.float32)
  N = len(R)
  assert N == 32000
  P = util.momenta(N, 3, kT=1.0, seed=0)
  dt, steps = 5e-3, 10_000
  d, s = jmd.space.periodic(L)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(
      d, L, r_onset=2.0, r_cutoff=2.5, dr_threshold=0.3, capacity_multiplier=1.6,
      format=jmd.partition.NeighborListFormat[fmt])
  Rd = _dev(R)
  nbrs = nf.allocate(Rd)
  init, step = jmd.simulate.nve(efn, s, dt)
  st = init(0, Rd, kT=1.0, momenta=_dev(P), neighbor=nbrs)
  KE = lambda st: float(jmd.quantity.kinetic_energy(momentum=st.momentum, mass=st.mass))
  E0 = float(efn(st.position, neighbor=nbrs)) + KE(st)

  def body(i, carry):
    st, nb = carry
    nb = nb.update(st.position)
    return step(st, neighbor=nb), nb
  b0 = nbrs._ws.state_host()[4]
  done = 0
  graph = None
  while done < steps:                     # the reference's loop shape: blocks + overflow check
    new_st, new_nb = jmd.lax.fori_loop(0, 1000, body, (st, nbrs), unroll=50, graph=graph)
    graph = jmd.lax.fori_loop.last
    if bool(new_nb.did_buffer_overflow):
      nbrs = nf.allocate(st.position)
      graph = None
      continue
    st, nbrs = new_st, new_nb
    done += 1000
  rebuilds = nbrs._ws.state_host()[4] - b0
  E1 = float(efn(st.position, neighbor=nbrs)) + KE(st)
  drift = (E1 - E0) / N
  drift_c, rebuilds_c = _cport_drift(R, L, P, dt, steps)
  bound = 2e-3
  assert abs(drift) < bound, drift
  assert abs(drift_c) < bound, drift_c
  assert abs(drift - drift_c) < bound
  # same physics, same skin rule: the rebuild cadence agrees
  assert abs(rebuilds - rebuilds_c) < 0.1 * rebuilds_c + 5, (rebuilds, rebuilds_c)
  # momentum is conserved by the full-list, atomics-free kernel
  assert float(st.momentum.sum(0, dtype=torch.float64).abs().max()) < 5e-2


# ---- N = 1M neighbour sets vs the C port ----------------------------------------------------

@pytest.mark.parametrize('dense', [True, False])
def test_gpu_neighbor_sets_match_c_port_1m(dense):
  jmd = _jmd()
  import bench
  R, box = bench.fcc((63, 63, 63))
  rng = np.random.default_rng(11)
  L = box[0]
  R = np.mod(R + rng.normal(0, 0.07, R.shape).astype(np.float32), L).astype(np.float32)
  N = len(R)
  assert N == 1_000_188
  sysc = cport.LJSystem(R, L, dense=dense)
  rows = sysc.rows()
  max_row = sysc.max_row
  sysc.close()
  d, _ = jmd.space.periodic(L)
  fmt = jmd.partition.Dense if dense else jmd.partition.Sparse
  nf = jmd.partition.neighbor_list(d, L, np.float32(2.5), np.float32(0.3), format=fmt)
  nbrs = nf.allocate(_dev(R))
  assert int(nbrs.error.code) == 0
  idx = nbrs.idx.cpu().numpy()
  if dense:
    assert nbrs.max_occupancy == int(max_row * 1.25)
    m = idx.shape[1]
    np.testing.assert_array_equal(np.sort(idx, -1), np.sort(rows[:, :m], -1))
  else:
    # (sender, receiver) pairs: senders ascending in the export, receivers sorted per sender
    total = int((idx[0] < N).sum())
    cnt_c = (rows < N).sum(1)
    assert total == int(cnt_c.sum())
    senders = idx[1, :total]
    np.testing.assert_array_equal(np.bincount(senders, minlength=N), cnt_c)
    starts = np.concatenate([[0], np.cumsum(cnt_c)])[:-1]
    # sort receivers inside every sender segment via a composite key
    key_g = senders.astype(np.int64) * N + idx[0, :total]
    key_g.sort()
    rs = np.sort(rows, -1)                      # valid entries first (padding is N)
    mk = rs < N
    key_c = (np.broadcast_to(np.arange(N, dtype=np.int64)[:, None], rs.shape)[mk] * N + rs[mk])
    np.testing.assert_array_equal(key_g, key_c)
    del starts


# ---- config 1: 2-D, f64, OrderedSparse NVE; 2-D FIRE ------------------------------------------

def test_2d_f64_ordered_sparse_nve_matches_oracle():
  """examples/nve_neighbor_list.py:90-201: square lattice, spacing 1.25, LJ defaults
  (r_onset 2, r_cutoff 2.5, dr_threshold 0.5), f64, OrderedSparse, dt = 1e-3."""
  jmd = _jmd()
  Nx, spacing = 30, np.float32(1.25)
  side = Nx * spacing
  R = np.stack([np.array(r) for r in np.ndindex(Nx, Nx)]).astype(np.float64) * float(spacing)
  rng = np.random.default_rng(0)
  R = np.mod(R + rng.normal(0, 0.02, R.shape), float(side))
  N = len(R)
  P = util.momenta(N, 2, kT=0.5, seed=1, dtype=np.float64)
  d_o, s_o = ospace.periodic(side)
  nf_o = opart.neighbor_list(d_o, side, np.float32(2.5), np.float32(0.5), format=opart.OrderedSparse)
  pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
  holder = {'nb': nf_o.allocate(R)}

  def f_o(Rx):
    holder['nb'] = holder['nb'].update(Rx)
    return oenergy.pair_neighbor_list_energy(pot, d_o, Rx, holder['nb'], want_grads=True,
                                             sigma=np.float64(1.0), epsilon=np.float64(1.0))[1]
  init_o, step_o = osim.nve(f_o, s_o, 1e-3)
  st_o = init_o(R, P, mass=np.float64(1.0))
  d_g, s_g = jmd.space.periodic(side)
  nf_g, efn = jmd.energy.lennard_jones_neighbor_list(d_g, side, format=jmd.partition.OrderedSparse)
  Rd = _dev(R)
  nbrs = nf_g.allocate(Rd)
  np.testing.assert_array_equal(util.sparse_pairs(nbrs.idx.cpu().numpy(), N),
                                util.sparse_pairs(holder['nb'].idx, N))
  init_g, step_g = jmd.simulate.nve(efn, s_g, 1e-3)
  st_g = init_g(0, Rd, kT=0.5, momenta=_dev(P), neighbor=nbrs)
  assert st_g.position.dtype == torch.float64
  b0 = nbrs._ws.state_host()[4]
  for _ in range(400):
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
  assert not bool(nbrs.did_buffer_overflow)
  assert nbrs._ws.state_host()[4] > b0            # crossed a rebuild
  dR = st_g.position.cpu().numpy() - st_o.position
  dR -= np.round(dR / float(side)) * float(side)
  assert np.abs(dR).max() < 1e-9
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum, atol=1e-8, rtol=0)
  # the list the GPU holds at the end is the oracle's list
  np.testing.assert_array_equal(util.sparse_pairs(nbrs.idx.cpu().numpy(), N),
                                util.sparse_pairs(holder['nb'].idx, N))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_2d_fire_descent_matches_oracle(dtype):
  """BASELINE config 3 semantics in 2-D: bidisperse soft spheres (sigma table of
  examples/fire_minimization.py), FIRE defaults, vs the oracle step by step."""
  jmd = _jmd()
  N = 1024
  rng = np.random.default_rng(3)
  L = np.float32(np.sqrt(N / 0.9))
  # a jittered square lattice: overlapping discs, but a smooth descent (uniformly random
  # discs start so far up the landscape that 1-ulp force differences decide F.P signs)
  g = np.stack([np.array(r) for r in np.ndindex(32, 32)]).astype(np.float64) * (float(L) / 32)
  R = np.mod(g + rng.normal(0, 0.12, g.shape), float(L)).astype(dtype)
  sp = (np.arange(N) % 2).astype(np.int32)
  sigma = np.array([[1.0, 1.2], [1.2, 1.4]], np.float32)
  d_o, s_o = ospace.periodic(L)
  # head-room: the packing rearranges during the descent (both sides get the same rule)
  nf_o = opart.neighbor_list(d_o, L, np.float32(1.4), np.float32(0.2), format=opart.OrderedSparse,
                             capacity_multiplier=2.0)
  pot = oenergy.PairPotential('soft_sphere')
  holder = {'nb': nf_o.allocate(R)}
  params = dict(sigma=sigma, epsilon=np.float32(1.0), alpha=np.float32(2.0))

  def f_o(Rx):
    # the list is the one update()d on the positions at the START of the step, exactly
    # like the reference loop (FIRE's first steps move atoms further than the skin)
    return oenergy.pair_neighbor_list_energy(pot, d_o, Rx, holder['nb'], species=sp,
                                             want_grads=True, **params)[1]
  init_o, step_o = osim.fire_descent(f_o, s_o)
  st_o = init_o(R, mass=dtype(1.0))
  d_g, s_g = jmd.space.periodic(L)
  nf_g, efn = jmd.energy.soft_sphere_neighbor_list(d_g, L, species=_dev(sp), sigma=sigma,
                                                   capacity_multiplier=2.0)
  Rd = _dev(R)
  nbrs = nf_g.allocate(Rd)
  init_g, step_g = jmd.minimize.fire_descent(efn, s_g)
  st_g = init_g(Rd, neighbor=nbrs)
  E0 = float(efn(Rd, neighbor=nbrs))
  for i in range(120):
    holder['nb'] = holder['nb'].update(st_o.position)
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
    if i in (20, 60):
      assert not bool(nbrs.did_buffer_overflow) and not holder['nb'].did_buffer_overflow
      assert int(st_g.n_pos) == st_o.n_pos
      np.testing.assert_allclose(float(st_g.dt), st_o.dt, rtol=1e-5)
      np.testing.assert_allclose(float(st_g.alpha), st_o.alpha, rtol=1e-5)
      dR = st_g.position.cpu().numpy() - st_o.position
      dR -= np.round(dR / float(L)) * float(L)
      assert np.abs(dR).max() < (1e-6 if dtype == np.float64 else 2e-3)
  assert not bool(nbrs.did_buffer_overflow)
  assert float(efn(st_g.position, neighbor=nbrs)) < 0.5 * E0
  assert float(st_g.force.abs().max()) < 3 * np.abs(st_o.force).max() + 1e-3
