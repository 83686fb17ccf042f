"""GPU parity: fused force/energy kernels and integrators vs the CPU oracle.

Tolerances are the north star's: rtol 1e-5 in f32, 1e-10 in f64 (the reference
tests use 2e-5 / 1e-10, tests/energy_test.py:63).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import energy as oenergy
from oracle import partition as opart
from oracle import simulate as osim
from oracle import space as ospace
from tests import util

pytestmark = pytest.mark.gpu

FORMATS = ['Dense', 'Sparse', 'OrderedSparse']
with open(os.path.join(util.GOLDEN, 'goldens.json')) as f:
  G = json.load(f)


def _jmd():
  import jax_md_b200 as jmd
  return jmd


def _dev(x):
  return torch.as_tensor(x, device='cuda')


def _tol(dtype):
  return dict(rtol=1e-5, atol=1e-5) if dtype == np.float32 else dict(rtol=1e-10, atol=1e-10)


def _ftol(dtype, F):
  scale = float(np.abs(F).max())
  if dtype == np.float32:
    return dict(rtol=1e-5, atol=1e-5 * scale)
  return dict(rtol=1e-10, atol=1e-10 * scale)


def _pair_setup(kind, dtype, fmt, n=6, species=False):
  jmd = _jmd()
  R, L = util.fcc(n, rho=0.8442, dtype=dtype)
  R = util.jitter(R, L, 0.06)
  N = len(R)
  d_o, s_o = ospace.periodic(L)
  d_g, s_g = jmd.space.periodic(L)
  F = jmd.partition.NeighborListFormat[fmt]
  sp = None
  kw_o, kw_g = {}, {}
  if species:
    sp = (np.arange(N) % 2).astype(np.int32)
  if kind == 'lj':
    sigma = np.array([[1.0, 1.05], [1.05, 1.1]], np.float32) if species else np.float32(1.0)
    eps = np.array([[1.0, 0.8], [0.8, 0.6]], np.float32) if species else np.float32(1.0)
    nf_g, e_g = jmd.energy.lennard_jones_neighbor_list(
        d_g, L, species=None if sp is None else _dev(sp), sigma=sigma,
        epsilon=eps, dr_threshold=0.3, format=F)
    smax = np.float32(np.max(sigma))
    pot = oenergy.PairPotential('lj', np.float32(2.0) * smax, np.float32(2.5) * smax)
    nf_o = opart.neighbor_list(d_o, L, np.float32(2.5) * smax, np.float32(0.3),
                               format=opart.Format[fmt])
    params = dict(sigma=sigma, epsilon=eps)
  elif kind == 'morse':
    nf_g, e_g = jmd.energy.morse_neighbor_list(d_g, L, sigma=1.1, epsilon=2.0,
                                               alpha=4.0, dr_threshold=0.3, format=F)
    pot = oenergy.PairPotential('morse', np.float32(2.0), np.float32(2.5))
    nf_o = opart.neighbor_list(d_o, L, np.float32(2.5), np.float32(0.3),
                               format=opart.Format[fmt])
    params = dict(sigma=np.float32(1.1), epsilon=np.float32(2.0), alpha=np.float32(4.0))
  else:
    sigma = np.array([[1.0, 1.2], [1.2, 1.4]], np.float32) if species else np.float32(1.3)
    nf_g, e_g = jmd.energy.soft_sphere_neighbor_list(
        d_g, L, species=None if sp is None else _dev(sp), sigma=sigma,
        dr_threshold=0.2, format=F)
    pot = oenergy.PairPotential('soft_sphere')
    nf_o = opart.neighbor_list(d_o, L, np.float32(np.max(sigma)), np.float32(0.2),
                               format=opart.Format[fmt])
    params = dict(sigma=sigma, epsilon=np.float32(1.0), alpha=np.float32(2.0))
  return dict(R=R, L=L, d_o=d_o, s_o=s_o, s_g=s_g, nf_o=nf_o, nf_g=nf_g, e_g=e_g,
              pot=pot, params=params, species=sp)


@pytest.mark.parametrize('kind', ['lj', 'soft_sphere', 'morse'])
@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_pair_energy_force_param_grads(kind, fmt, dtype):
  jmd = _jmd()
  s = _pair_setup(kind, dtype, fmt)
  R = s['R']
  nb_o = s['nf_o'].allocate(R)
  Rd = _dev(R)
  nb_g = s['nf_g'].allocate(Rd)
  np.testing.assert_array_equal(np.sort(nb_g.idx.cpu().numpy(), -1), np.sort(nb_o.idx, -1)) \
      if fmt == 'Dense' else np.testing.assert_array_equal(
          util.sparse_pairs(nb_g.idx.cpu().numpy(), len(R)), util.sparse_pairs(nb_o.idx, len(R)))
  E_o, F_o, dp_o = oenergy.pair_neighbor_list_energy(
      s['pot'], s['d_o'], R.astype(np.float64), nb_o, want_grads=True,
      **{k: np.float64(v) for k, v in s['params'].items()})
  E_g = s['e_g'](Rd, neighbor=nb_g)
  assert E_g.dtype == Rd.dtype and E_g.ndim == 0
  np.testing.assert_allclose(float(E_g), E_o, **_tol(dtype))
  F_g = jmd.quantity.force(s['e_g'])(Rd, neighbor=nb_g)
  np.testing.assert_allclose(F_g.cpu().numpy(), F_o, **_ftol(dtype, F_o))
  # autograd route == jax.grad(energy_fn): positions and (sigma, epsilon)
  Rg = Rd.clone().requires_grad_(True)
  sig = torch.tensor(float(s['params']['sigma']), dtype=Rd.dtype, device='cuda',
                     requires_grad=True)
  eps = torch.tensor(float(s['params']['epsilon']), dtype=Rd.dtype, device='cuda',
                     requires_grad=True)
  E = s['e_g'](Rg, neighbor=nb_g, sigma=sig, epsilon=eps)
  E.backward()
  np.testing.assert_allclose((-Rg.grad).cpu().numpy(), F_o, **_ftol(dtype, F_o))
  t = dict(rtol=2e-5, atol=1e-4) if dtype == np.float32 else dict(rtol=1e-9, atol=1e-9)
  np.testing.assert_allclose(float(sig.grad), dp_o['sigma'], **t)
  np.testing.assert_allclose(float(eps.grad), dp_o['epsilon'], **t)


@pytest.mark.parametrize('kind', ['lj', 'soft_sphere'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_species_tables_and_per_particle(kind, dtype):
  jmd = _jmd()
  s = _pair_setup(kind, dtype, 'Sparse', species=True)
  R = s['R']
  nb_o = s['nf_o'].allocate(R)
  Rd = _dev(R)
  nb_g = s['nf_g'].allocate(Rd)
  p64 = {k: np.asarray(v, np.float64) for k, v in s['params'].items()}
  E_o, F_o, dp_o = oenergy.pair_neighbor_list_energy(
      s['pot'], s['d_o'], R.astype(np.float64), nb_o, species=s['species'],
      want_grads=True, **p64)
  E_g = s['e_g'](Rd, nb_g)          # positional neighbor, as tests/energy_test.py:596
  np.testing.assert_allclose(float(E_g), E_o, **_tol(dtype))
  F_g = jmd.quantity.force(s['e_g'])(Rd, neighbor=nb_g)
  np.testing.assert_allclose(F_g.cpu().numpy(), F_o, **_ftol(dtype, F_o))
  # table gradients
  sig = _dev(np.asarray(s['params']['sigma'], dtype)).requires_grad_(True)
  E = s['e_g'](Rd, neighbor=nb_g, sigma=sig)
  E.backward()
  t = dict(rtol=5e-5, atol=1e-3) if dtype == np.float32 else dict(rtol=1e-9, atol=1e-9)
  np.testing.assert_allclose(sig.grad.cpu().numpy(), dp_o['sigma'], **t)
  # ... and they are bitwise reproducible: per-atom rows folded by a GEMM, no atomics
  # (jmd_pair_t.dparam_rows)
  g0 = sig.grad.clone()
  for _ in range(3):
    sig.grad = None
    s['e_g'](Rd, neighbor=nb_g, sigma=sig).backward()
    assert torch.equal(sig.grad, g0)
  eps_t = _dev(np.asarray(s['params']['epsilon'], dtype)).requires_grad_(True)
  s['e_g'](Rd, neighbor=nb_g, epsilon=eps_t).backward()
  np.testing.assert_allclose(eps_t.grad.cpu().numpy(), dp_o['epsilon'], **t)
  # per-particle energies (reduce_axis=(1,), smap.py:958-977)
  if kind == 'lj':
    d_g, _ = jmd.space.periodic(s['L'])
    _, e_pp = jmd.energy.lennard_jones_neighbor_list(
        d_g, s['L'], species=_dev(s['species']), sigma=s['params']['sigma'],
        epsilon=s['params']['epsilon'], dr_threshold=0.3, per_particle=True,
        format=jmd.partition.Sparse)
    Ea_o = oenergy.pair_neighbor_list_energy(
        s['pot'], s['d_o'], R.astype(np.float64), nb_o, species=s['species'],
        per_particle=True, **p64)
    Ea_g = e_pp(Rd, neighbor=nb_g)
    np.testing.assert_allclose(Ea_g.cpu().numpy(), Ea_o,
                               rtol=1e-5 if dtype == np.float32 else 1e-10,
                               atol=1e-5 if dtype == np.float32 else 1e-10)


@pytest.mark.parametrize('fmt', FORMATS)
def test_jammed_soft_sphere_golden(fmt):
  """tests/data/simulation_test_state.npy: E = 0.45247561922261154 (2-D, f64)."""
  jmd = _jmd()
  s = np.load(os.path.join(util.GOLDEN, 'jammed_state.npz'))
  R = _dev(s['real_position'])
  L = float(s['box'][0, 0])
  d, _ = jmd.space.periodic(L)
  nf, efn = jmd.energy.soft_sphere_neighbor_list(
      d, L, species=_dev(s['species']), sigma=s['sigma'],
      format=jmd.partition.NeighborListFormat[fmt])
  nbrs = nf.allocate(R)
  np.testing.assert_allclose(float(efn(R, neighbor=nbrs)), G['jammed_energy'],
                             rtol=1e-10)
  # tests/quantity_test.py:134-150: P = 0.06307342050945483 of the same state, here
  # from the virial the fused force kernel accumulates (quantity.pressure)
  P = jmd.quantity.pressure(efn, R, L, neighbor=nbrs)
  np.testing.assert_allclose(float(P), G['jammed_pressure'], rtol=1e-9)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('n', [3, 4])
def test_stillinger_weber_golden_and_forces(dtype, n):
  """tests/energy_test.py:429-466: -4.336503155764325 eV/atom on diamond Si;
  forces on a perturbed lattice vs the oracle."""
  jmd = _jmd()
  R, L = util.diamond(n, a=G['sw_lattice_constant'], dtype=dtype)
  d_g, _ = jmd.space.periodic(L)
  nf, efn = jmd.energy.stillinger_weber_neighbor_list(d_g, L)
  Rd = _dev(R)
  nbrs = nf.allocate(Rd)
  E = float(efn(Rd, neighbor=nbrs)) / len(R)
  np.testing.assert_allclose(E, G['sw_diamond_energy_per_atom'],
                             rtol=2e-5 if dtype == np.float32 else 1e-10)
  Rp = util.jitter(R, L, 0.08, seed=2)
  d_o, _ = ospace.periodic(L)
  nf_o = opart.neighbor_list(d_o, L, 3.77118, 0.5, format=opart.Dense)
  nb_o = nf_o.allocate(Rp)
  E_o, F_o = oenergy.stillinger_weber_energy(d_o, Rp.astype(np.float64), nb_o,
                                             want_force=True)
  Rpd = _dev(Rp)
  nbrs = nf.allocate(Rpd)
  np.testing.assert_array_equal(np.sort(nbrs.idx.cpu().numpy(), -1), np.sort(nb_o.idx, -1))
  E_g = float(efn(Rpd, neighbor=nbrs))
  np.testing.assert_allclose(E_g, E_o, rtol=2e-5 if dtype == np.float32 else 1e-10)
  F_g = jmd.quantity.force(efn)(Rpd, neighbor=nbrs)
  np.testing.assert_allclose(F_g.cpu().numpy(), F_o, **_ftol(dtype, F_o))


# -- integrators -----------------------------------------------------------------

def _oracle_force(s, nbrs_holder):
  def f(Rx):
    nb = nbrs_holder['nb'].update(Rx)
    nbrs_holder['nb'] = nb
    return oenergy.pair_neighbor_list_energy(
        s['pot'], s['d_o'], Rx, nb, want_grads=True, **s['params'])[1]
  return f


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('fmt', ['Dense', 'OrderedSparse'])
def test_nve_trajectory_matches_oracle(dtype, fmt):
  """tests/simulate_test.py:203-247 analogue: NL trajectory vs the oracle's."""
  jmd = _jmd()
  s = _pair_setup('lj', dtype, fmt, n=5)
  R = s['R']
  P = util.momenta(len(R), 3, kT=1.0, dtype=dtype)
  holder = {'nb': s['nf_o'].allocate(R)}
  init_o, step_o = osim.nve(_oracle_force(s, holder), s['s_o'], 1e-3)
  st_o = init_o(R, P, mass=dtype(1.0))
  init_g, step_g = jmd.simulate.nve(s['e_g'], s['s_g'], 1e-3)
  Rd = _dev(R)
  nbrs = s['nf_g'].allocate(Rd)
  st_g = init_g(0, Rd, kT=1.0, momenta=_dev(P), neighbor=nbrs)
  np.testing.assert_allclose(st_g.force.cpu().numpy(), st_o.force,
                             **_ftol(dtype, st_o.force))
  steps = 200
  for _ in range(steps):
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
  assert not bool(nbrs.did_buffer_overflow)
  tol = 5e-3 if dtype == np.float32 else 5e-9      # simulate_test.py:243-246 uses 5e-3 / 5e-12 over 2000
  dR = st_g.position.cpu().numpy() - st_o.position
  dR -= np.round(dR / s['L']) * s['L']
  assert np.abs(dR).max() < tol
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum, atol=tol * 10, rtol=0)
  assert st_g.position.dtype == Rd.dtype


def test_nve_energy_conservation_and_rebuilds():
  """NVE drift over 2000 steps at T*=1 with rebuilds through the fused update
  path (f32): |dE|/N below 2e-4 (stated bound), no overflow."""
  jmd = _jmd()
  R, L = util.fcc(10, dtype=np.float32)
  d, s = jmd.space.periodic(L)
  # the perfect lattice under-estimates the liquid's cell occupancy: give the
  # buffers head-room (the reference needs the same, partition.py:867-870)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(d, L, dr_threshold=0.3,
                                                   capacity_multiplier=1.6)
  Rd = _dev(R)
  nbrs = nf.allocate(Rd)
  init, step = jmd.simulate.nve(efn, s, 5e-3)
  st = init(0, Rd, kT=1.0, momenta=_dev(util.momenta(len(R), 3, 1.0)), neighbor=nbrs)
  KE = lambda st: float(jmd.quantity.kinetic_energy(momentum=st.momentum, mass=st.mass))
  E0 = float(efn(st.position, neighbor=nbrs)) + KE(st)
  rebuilds = 0
  # the reference's own loop shape (partition.py:840-854): blocks of steps,
  # re-allocate from the last good state when a buffer overflowed.
  for _ in range(20):
    b0 = nbrs._ws.state_host()[4]
    new_st, new_nbrs = st, nbrs
    for _ in range(100):
      new_nbrs = new_nbrs.update(new_st.position)
      new_st = step(new_st, neighbor=new_nbrs)
    if bool(new_nbrs.did_buffer_overflow):
      nbrs = nf.allocate(st.position)
      new_st, new_nbrs = st, nbrs
      for _ in range(100):
        new_nbrs = new_nbrs.update(new_st.position)
        new_st = step(new_st, neighbor=new_nbrs)
      assert not bool(new_nbrs.did_buffer_overflow)
    rebuilds += new_nbrs._ws.state_host()[4] - b0
    st, nbrs = new_st, new_nbrs
  E1 = float(efn(st.position, neighbor=nbrs)) + KE(st)
  assert rebuilds > 20                               # it did rebuild
  assert abs(E1 - E0) / len(R) < 2e-4


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('sy', [1, 3, 5, 7])
def test_nvt_nose_hoover_matches_oracle(dtype, sy):
  """tests/simulate_test.py:264-350: chain variables and momenta vs the oracle,
  and the NHC invariant is conserved."""
  jmd = _jmd()
  s = _pair_setup('lj', dtype, 'Dense', n=5)
  R = s['R']
  kT = 0.9
  P = util.momenta(len(R), 3, kT=kT, dtype=dtype)
  holder = {'nb': s['nf_o'].allocate(R)}
  init_o, step_o = osim.nvt_nose_hoover(_oracle_force(s, holder), s['s_o'], 1e-3,
                                        kT, chain_length=3, sy_steps=sy)
  st_o = init_o(R, P, mass=dtype(1.0))
  init_g, step_g = jmd.simulate.nvt_nose_hoover(s['e_g'], s['s_g'], 1e-3, kT,
                                                chain_length=3, sy_steps=sy)
  Rd = _dev(R)
  nbrs = s['nf_g'].allocate(Rd)
  st_g = init_g(0, Rd, momenta=_dev(P), neighbor=nbrs)
  H0 = float(jmd.simulate.nvt_nose_hoover_invariant(s['e_g'], st_g, kT, neighbor=nbrs))
  for _ in range(100):
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
  H1 = float(jmd.simulate.nvt_nose_hoover_invariant(s['e_g'], st_g, kT, neighbor=nbrs))
  rt = 2e-3 if dtype == np.float32 else 1e-7
  np.testing.assert_allclose(st_g.chain.momentum.cpu().numpy(), st_o.chain.momentum,
                             rtol=rt, atol=rt)
  np.testing.assert_allclose(st_g.chain.position.cpu().numpy(), st_o.chain.position,
                             rtol=rt, atol=rt)
  np.testing.assert_allclose(float(st_g.chain.kinetic_energy),
                             float(st_o.chain.kinetic_energy), rtol=rt)
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_o.momentum, atol=rt, rtol=0)
  # simulate_test.py:264-350 uses rtol 5e-4 (f32) / 1e-6 (f64) on its own
  # system; the drift here is the integrator's (dt=1e-3, T*=0.9), identical in
  # the oracle, so bound it by 1e-5.
  assert abs(H1 - H0) < (5e-4 if dtype == np.float32 else 1e-5) * abs(H0)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_fire_descent_matches_oracle(dtype):
  """tests/minimize_test.py:111 + step-by-step schedule parity with the oracle."""
  jmd = _jmd()
  s = _pair_setup('soft_sphere', dtype, 'OrderedSparse', n=5, species=True)
  R = s['R']

  def f_o(Rx):
    nb = holder['nb'].update(Rx)
    holder['nb'] = nb
    return oenergy.pair_neighbor_list_energy(
        s['pot'], s['d_o'], Rx, nb, species=s['species'], want_grads=True,
        **s['params'])[1]
  holder = {'nb': s['nf_o'].allocate(R)}
  init_o, step_o = osim.fire_descent(f_o, s['s_o'])
  st_o = init_o(R, mass=dtype(1.0))
  init_g, step_g = jmd.minimize.fire_descent(s['e_g'], s['s_g'])
  Rd = _dev(R)
  nbrs = s['nf_g'].allocate(Rd)
  st_g = init_g(Rd, neighbor=nbrs)
  f0 = float(st_g.force.abs().max())
  for i in range(120):
    st_o = step_o(st_o)
    nbrs = nbrs.update(st_g.position)
    st_g = step_g(st_g, neighbor=nbrs)
    if i == 30:
      assert int(st_g.n_pos) == st_o.n_pos
      np.testing.assert_allclose(float(st_g.dt), st_o.dt, rtol=1e-5)
      np.testing.assert_allclose(float(st_g.alpha), st_o.alpha, rtol=1e-5)
  assert float(st_g.force.abs().max()) < 0.25 * f0
  np.testing.assert_allclose(float(st_g.force.abs().max()), np.abs(st_o.force).max(),
                             rtol=1e-2 if dtype == np.float32 else 1e-6)
  if dtype == np.float64:
    np.testing.assert_allclose(float(st_g.dt), st_o.dt, rtol=1e-9)
    dR = st_g.position.cpu().numpy() - st_o.position
    dR -= np.round(dR / s['L']) * s['L']
    assert np.abs(dR).max() < 1e-7


def test_generic_force_fn_path():
  """simulate.nve over a user force function (not fused): the integrator
  kernels bracket a Python callable (quantity.canonicalize_force)."""
  jmd = _jmd()
  _, s = jmd.space.periodic(10.0)
  k = 2.0
  R = _dev(np.random.default_rng(0).random((256, 3)).astype(np.float64) * 2 + 4)
  efn = lambda Rx, **kw: 0.5 * k * ((Rx - 5.0) ** 2).sum()
  init, step = jmd.simulate.nve(efn, s, 1e-2)
  st = init(1, R, kT=0.0, momenta=torch.zeros_like(R))
  E0 = float(efn(st.position))
  for _ in range(100):
    st = step(st)
  E1 = float(efn(st.position)) + float(jmd.quantity.kinetic_energy(
      momentum=st.momentum, mass=st.mass))
  assert abs(E1 - E0) < 1e-3 * E0


@pytest.mark.parametrize('kind', ['nve', 'nvt', 'fire'])
def test_fori_loop_cuda_graph_matches_eager(kind):
  """lax.fori_loop (CUDA-graph replay of the step body, rebuild decision on the
  device) follows the eager loop; it is the analogue of jit(lax.fori_loop)."""
  jmd = _jmd()
  dtype = np.float64
  s = _pair_setup('lj' if kind != 'fire' else 'soft_sphere', dtype, 'Dense', n=6)
  R = _dev(s['R'])
  P = _dev(util.momenta(len(s['R']), 3, kT=1.0, dtype=dtype))
  if kind == 'nve':
    init, step = jmd.simulate.nve(s['e_g'], s['s_g'], 2e-3)
    mk = lambda nb: init(0, R, kT=1.0, momenta=P, neighbor=nb)
  elif kind == 'nvt':
    init, step = jmd.simulate.nvt_nose_hoover(s['e_g'], s['s_g'], 2e-3, 0.8, chain_length=3)
    mk = lambda nb: init(0, R, momenta=P, neighbor=nb)
  else:
    init, step = jmd.minimize.fire_descent(s['e_g'], s['s_g'])
    mk = lambda nb: init(R, neighbor=nb)

  def body(i, carry):
    st, nb = carry
    nb = nb.update(st.position)
    return step(st, neighbor=nb), nb
  steps = 120
  nb_e = s['nf_g'].allocate(R)
  st_e = mk(nb_e)
  for i in range(steps):
    st_e, nb_e = body(i, (st_e, nb_e))
  nb_g = s['nf_g'].allocate(R)
  st_g = mk(nb_g)
  b0 = nb_g._ws.state_host()[4]
  st_g, nb_g = jmd.lax.fori_loop(0, steps, body, (st_g, nb_g), unroll=20)
  assert not bool(nb_g.did_buffer_overflow)
  if kind != 'fire':
    assert nb_g._ws.state_host()[4] > b0          # rebuilt inside the graph
  dR = (st_g.position - st_e.position).cpu().numpy()
  dR -= np.round(dR / s['L']) * s['L']
  assert np.abs(dR).max() < 1e-8
  np.testing.assert_allclose(st_g.momentum.cpu().numpy(), st_e.momentum.cpu().numpy(), atol=1e-7)
  if kind == 'nvt':
    np.testing.assert_allclose(st_g.chain.momentum.cpu().numpy(), st_e.chain.momentum.cpu().numpy(),
                               rtol=1e-7, atol=1e-9)
  if kind == 'fire':
    assert int(st_g.n_pos) == int(st_e.n_pos)
    np.testing.assert_allclose(float(st_g.dt), float(st_e.dt), rtol=1e-12)


@pytest.mark.parametrize('kind', ['lj', 'soft_sphere'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_staged_force_kernel_bitwise_equals_direct(kind, dtype):
  """The shared-memory staged force kernel (16-bit staging rows) evaluates the
  same pairs in the same order as the global-gather kernel: forces and the fused
  half kick (hence trajectories) are bitwise equal, energies to summation order.  N=13500: 1-2 row segments per block,
  x wrap pieces, partially filled last block."""
  jmd = _jmd()
  R, L = util.fcc(15, dtype=dtype)
  R = util.jitter(R, L, 0.05)
  N = len(R)
  d_g, s_g = jmd.space.periodic(L)
  sp = _dev((np.arange(N) % 2).astype(np.int32))
  outs = []
  for stage in (True, False):
    if kind == 'lj':
      nf, efn = jmd.energy.lennard_jones_neighbor_list(
          d_g, L, dr_threshold=0.3, format=jmd.partition.OrderedSparse, stage_positions=stage)
    else:
      nf, efn = jmd.energy.soft_sphere_neighbor_list(
          d_g, L, species=sp, sigma=np.array([[1.0, 1.2], [1.2, 1.4]], np.float32),
          dr_threshold=0.2, format=jmd.partition.Dense, stage_positions=stage)
    Rd = _dev(R)
    nb = nf.allocate(Rd)
    ws = nb._ws
    assert ws.c.staged == (1 if stage else 0)
    if stage:
      modes = ws.t['blk_table'][:(N + 255) // 256, 0]
      cap = 57344 // (16 if dtype == np.float32 else 32)
      if dtype == np.float32 and kind == 'lj':   # short rows (9 cells): a few blocks span too many rows
        assert int(modes.sum()) >= 0.8 * modes.numel(), 'most blocks should be staged here'
      assert int((ws.t['blk_table'][:(N + 255) // 256, 1] * modes).max()) <= cap
    E = efn(Rd, neighbor=nb)
    F = jmd.quantity.force(efn)(Rd, neighbor=nb)
    init, step = jmd.simulate.nve(efn, s_g, 5e-3)
    st = init(0, Rd, kT=1.0, momenta=_dev(util.momenta(N, 3, dtype=dtype)), neighbor=nb)
    for _ in range(25):                       # crosses at least one rebuild
      nb = nb.update(st.position)
      st = step(st, neighbor=nb)
    outs.append((E, F, st.position.clone(), st.momentum.clone(), nb.idx.clone()))
  # per-atom results are bitwise equal; the total energy is a block-wise f64 sum and
  # the two kernels use different block sizes (256 staged / 128 gather)
  np.testing.assert_allclose(float(outs[0][0]), float(outs[1][0]), rtol=1e-13)
  for a, b in zip(outs[0][1:], outs[1][1:]):
    assert torch.equal(a, b)


@pytest.mark.parametrize('fmt', ['Dense', 'OrderedSparse'])
def test_fused_skin_predicate_same_rebuild_decisions(fmt):
  """The skin predicate evaluated inside the drift kernel (and reused by
  update() when it is handed that very tensor) takes the same rebuild
  decisions as update()'s own pass: same rebuild steps, bitwise-equal states."""
  jmd = _jmd()
  R, L = util.fcc(10, dtype=np.float32)
  R = util.jitter(R, L, 0.05)
  N = len(R)
  d_g, s_g = jmd.space.periodic(L)
  out = []
  for clone in (False, True):
    nf, efn = jmd.energy.lennard_jones_neighbor_list(
        d_g, L, dr_threshold=0.3, format=jmd.partition.NeighborListFormat[fmt])
    init, step = jmd.simulate.nve(efn, s_g, 5e-3)
    Rd = _dev(R)
    nb = nf.allocate(Rd)
    st = init(0, Rd, kT=1.5, momenta=_dev(util.momenta(N, 3, kT=1.5)), neighbor=nb)
    builds = []
    for i in range(60):
      pos = st.position.clone() if clone else st.position     # a clone is never "the drift's tensor"
      nb = nb.update(pos)
      builds.append(nb._ws.state_host()[4])
      st = step(st, neighbor=nb)
      if i == 30 and not clone:
        st.position.add_(0.0)      # in-place touch: version bump -> update() must not trust the flags
    out.append((builds, st.position.clone(), st.momentum.clone(), nb.idx.clone()))
  assert out[0][0] == out[1][0]
  assert out[0][0][-1] - out[0][0][0] >= 3, 'the run should cross several rebuilds'
  for a, b in zip(out[0][1:], out[1][1:]):
    assert torch.equal(a, b)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_generic_pair_neighbor_list_matches_fused(fmt, dtype):
  """smap.pair_neighbor_list over an arbitrary Python fn: the
  torch-composed generic path over the CUDA-built list gives the same energy,
  forces and per-atom energies as the fused kernel for the same potential
  (scalar, per-atom and species-table parameters)."""
  jmd = _jmd()
  R, L = util.fcc(7, dtype=dtype)
  R = util.jitter(R, L, 0.06)
  N = len(R)
  d_g, s_g = jmd.space.periodic(L)
  F = jmd.partition.NeighborListFormat[fmt]
  Rd = _dev(R)
  sp = _dev((np.arange(N) % 2).astype(np.int32))
  sig_tab = _dev(np.array([[1.0, 1.05], [1.05, 1.1]], dtype))
  sig_atom = _dev((1.0 + 0.05 * (np.arange(N) % 3)).astype(dtype))

  def my_lj(dr, sigma=1.0, epsilon=1.0, **kw):           # a user lambda: no fused tag
    x6 = (sigma / dr) ** 6
    return torch.nan_to_num(4 * epsilon * (x6 * x6 - x6))

  cut = lambda f: jmd.energy.multiplicative_isotropic_cutoff(f, np.float32(2.0), np.float32(2.5))
  nf = jmd.partition.neighbor_list(d_g, L, np.float32(2.5), np.float32(0.3), format=F)
  nb = nf.allocate(Rd)
  tol = dict(rtol=1e-5, atol=1e-5) if dtype == np.float32 else dict(rtol=1e-9, atol=1e-9)
  cases = [dict(), dict(sigma=sig_atom), dict(sigma=sig_tab, species=sp)]
  for kw in cases:
    gen = jmd.smap.pair_neighbor_list(cut(my_lj), d_g, **kw)
    fus = jmd.smap.pair_neighbor_list(cut(jmd.energy.lennard_jones), d_g, **kw)
    assert isinstance(gen, jmd.smap.GenericPairNeighborListFn)
    assert not isinstance(fus, jmd.smap.GenericPairNeighborListFn)
    Eg, Ef = float(gen(Rd, neighbor=nb)), float(fus(Rd, neighbor=nb))
    np.testing.assert_allclose(Eg, Ef, rtol=tol['rtol'])
    Fg = jmd.quantity.force(gen)(Rd, neighbor=nb).cpu().numpy()
    Ff = jmd.quantity.force(fus)(Rd, neighbor=nb).cpu().numpy()
    np.testing.assert_allclose(Fg, Ff, **_ftol(dtype, Ff))
  if fmt != 'OrderedSparse':
    gen1 = jmd.smap.pair_neighbor_list(cut(my_lj), d_g, reduce_axis=(1,))
    fus1 = jmd.smap.pair_neighbor_list(cut(jmd.energy.lennard_jones), d_g, reduce_axis=(1,))
    np.testing.assert_allclose(gen1(Rd, neighbor=nb).cpu().numpy(), fus1(Rd, neighbor=nb).cpu().numpy(),
                               rtol=10 * tol['rtol'], atol=10 * tol['atol'])
  # and it drives the integrator through autograd forces
  gen = jmd.smap.pair_neighbor_list(cut(my_lj), d_g)
  init, step = jmd.simulate.nve(gen, s_g, 1e-3)
  st = init(0, Rd, kT=0.5, momenta=_dev(util.momenta(N, 3, kT=0.5, dtype=dtype)), neighbor=nb)
  E0 = float(gen(st.position, neighbor=nb)) + float(jmd.quantity.kinetic_energy(momentum=st.momentum))
  for _ in range(20):
    nb = nb.update(st.position)
    st = step(st, neighbor=nb)
  E1 = float(gen(st.position, neighbor=nb)) + float(jmd.quantity.kinetic_energy(momentum=st.momentum))
  assert abs(E1 - E0) < 2e-4 * abs(E0)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_stress_lammps_golden_and_generic_path(dtype):
  """tests/quantity_test.py:436-455: LAMMPS LJ (hard cutoff 2.5) energy per atom and
  stress tensor incl. the kinetic term; the fused kernel's virial, and the same
  numbers from the generic path (autograd through `perturbation=`)."""
  jmd = _jmd()
  s = np.load(os.path.join(util.GOLDEN, 'lammps_lj.npz'))
  box = np.float32(s['box'])
  R = _dev((s['R'] * box).astype(dtype))
  V = _dev(s['V'].astype(dtype))
  r = s['stress_row']
  C = np.array([[r[0], r[3], r[4]], [r[3], r[1], r[5]], [r[4], r[5], r[2]]])
  d, _ = jmd.space.periodic(box)
  # list cutoff 2.5 with no skin == the hard cutoff of the LAMMPS run
  nf = jmd.partition.neighbor_list(d, box, np.float32(2.5), np.float32(0.0), format=jmd.partition.Dense)
  nb = nf.allocate(R)
  fused = jmd.smap.pair_neighbor_list(jmd.energy.lennard_jones, d)
  generic = jmd.smap.pair_neighbor_list(lambda dr, **kw: jmd.energy.lennard_jones(dr), d)
  tol = 5e-5
  for efn in (fused, generic):
    np.testing.assert_allclose(float(efn(R, neighbor=nb)) / len(R), float(s['energy_per_atom']),
                               rtol=tol, atol=tol)
    S = jmd.quantity.stress(efn, R, box, velocity=V, neighbor=nb).cpu().numpy()
    np.testing.assert_allclose(S, C, rtol=tol, atol=tol)
  Pf = float(jmd.quantity.pressure(fused, R, box, neighbor=nb))
  Pg = float(jmd.quantity.pressure(generic, R, box, neighbor=nb))
  np.testing.assert_allclose(Pf, Pg, rtol=1e-5 if dtype == np.float32 else 1e-10)


@pytest.mark.parametrize('fmt', FORMATS)
def test_asymmetric_species_table_orientation(fmt):
  """smap.py:794-797: Dense looks up p[species[row], species[neighbour]], the sparse
  formats p[species[idx[0]], species[idx[1]]].  An asymmetric table tells them apart."""
  jmd = _jmd()
  R, L = util.fcc(6, dtype=np.float64)
  R = util.jitter(R, L, 0.06)
  N = len(R)
  sp = (np.arange(N) % 2).astype(np.int32)
  sig = np.array([[1.0, 0.9], [1.15, 1.05]])          # sigma[0,1] != sigma[1,0]
  d_o, _ = ospace.periodic(L)
  d_g, _ = jmd.space.periodic(L)
  nf_o = opart.neighbor_list(d_o, L, np.float32(1.4), np.float32(0.2), format=opart.Format[fmt])
  nb_o = nf_o.allocate(R)
  pot = oenergy.PairPotential('soft_sphere')
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(pot, d_o, R, nb_o, species=sp, want_grads=True,
                                                  sigma=sig, epsilon=np.float64(1.0), alpha=np.float64(2.0))
  Rd = _dev(R)
  nf_g = jmd.partition.neighbor_list(d_g, L, np.float32(1.4), np.float32(0.2),
                                     format=jmd.partition.NeighborListFormat[fmt])
  nb_g = nf_g.allocate(Rd)
  fused = jmd.smap.pair_neighbor_list(jmd.energy.soft_sphere, d_g, species=_dev(sp), sigma=_dev(sig))
  generic = jmd.smap.pair_neighbor_list(lambda dr, sigma=1.0, **kw: jmd.energy.soft_sphere(dr, sigma), d_g,
                                        species=_dev(sp), sigma=_dev(sig))
  np.testing.assert_allclose(float(generic(Rd, neighbor=nb_g)), E_o, rtol=1e-10)
  F = jmd.quantity.force(generic)(Rd, neighbor=nb_g).cpu().numpy()
  np.testing.assert_allclose(F, F_o, rtol=1e-8, atol=1e-9 * np.abs(F_o).max())
  # the fused kernel's per-row force needs symmetric tables: the factory notices the
  # asymmetric table and serves this energy through the generic path instead of guessing
  assert fused.always_generic
  np.testing.assert_allclose(float(fused(Rd, neighbor=nb_g)), E_o, rtol=1e-10)
  np.testing.assert_allclose(jmd.quantity.force(fused)(Rd, neighbor=nb_g).cpu().numpy(), F_o,
                             rtol=1e-8, atol=1e-9 * np.abs(F_o).max())
  sym = jmd.smap.pair_neighbor_list(jmd.energy.soft_sphere, d_g, species=_dev(sp),
                                    sigma=_dev(0.5 * (sig + sig.T)))
  E_s, F_s, _ = oenergy.pair_neighbor_list_energy(pot, d_o, R, nb_o, species=sp, want_grads=True,
                                                  sigma=0.5 * (sig + sig.T), epsilon=np.float64(1.0),
                                                  alpha=np.float64(2.0))
  np.testing.assert_allclose(float(sym(Rd, neighbor=nb_g)), E_s, rtol=1e-10)
  np.testing.assert_allclose(jmd.quantity.force(sym)(Rd, neighbor=nb_g).cpu().numpy(), F_s,
                             rtol=1e-8, atol=1e-9 * np.abs(F_s).max())
