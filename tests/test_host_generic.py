"""CPU tests of host-side logic that needs no device: the generic
`smap.pair_neighbor_list` path (torch ops) against the NumPy
oracle on oracle-built neighbour lists, `quantity.volume / pressure / stress`
through autograd against the reference's goldens, and the `NeighborList`
dataclass contract (replace / set / frozen / idx descriptor)."""
import dataclasses
import json
import os

import numpy as np
import pytest
import torch

from oracle import energy as oenergy
from oracle import partition as opart
from oracle import space as ospace
from tests import util

with open(os.path.join(util.GOLDEN, 'goldens.json')) as f:
  G = json.load(f)

FORMATS = ['Dense', 'Sparse', 'OrderedSparse']


def _wrap(nb_o, fmt):
  """An oracle list as a jax_md_b200 NeighborList (no device workspace)."""
  from jax_md_b200 import partition as P
  return P.NeighborList(torch.as_tensor(np.asarray(nb_o.idx)), torch.as_tensor(nb_o.reference_position),
                        P.PartitionError(torch.tensor(int(nb_o.error), dtype=torch.uint8)),
                        nb_o.cell_list_capacity, nb_o.max_occupancy, P.NeighborListFormat[fmt],
                        None, None, None, None)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('kind', ['lj', 'soft_sphere'])
def test_generic_pair_path_matches_oracle(fmt, kind):
  from jax_md_b200 import energy, quantity, smap, space
  R, L = util.fcc(5, dtype=np.float64)
  R = util.jitter(R, L, 0.06)
  N = len(R)
  sp = (np.arange(N) % 2).astype(np.int32)
  d_o, _ = ospace.periodic(L)
  d_t, _ = space.periodic(L)
  if kind == 'lj':
    sig = np.array([[1.0, 1.05], [1.05, 1.1]])
    pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
    rc, skin = np.float32(2.5), np.float32(0.3)
    fn = energy.multiplicative_isotropic_cutoff(
        lambda dr, sigma=1.0, epsilon=1.0, **kw: energy.lennard_jones(dr, sigma, epsilon),
        np.float32(2.0), np.float32(2.5))
    params = dict(sigma=sig, epsilon=np.float64(1.0))
  else:
    sig = np.array([[1.0, 1.2], [1.2, 1.4]])
    pot = oenergy.PairPotential('soft_sphere')
    rc, skin = np.float32(1.4), np.float32(0.2)
    fn = lambda dr, sigma=1.0, epsilon=1.0, alpha=2.0, **kw: energy.soft_sphere(dr, sigma, epsilon, alpha)
    params = dict(sigma=sig, epsilon=np.float64(1.0), alpha=np.float64(2.0))
  nb_o = opart.neighbor_list(d_o, L, rc, skin, format=opart.Format[fmt]).allocate(R)
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(pot, d_o, R, nb_o, species=sp, want_grads=True, **params)
  gen = smap.pair_neighbor_list(fn, d_t, species=torch.as_tensor(sp),
                                **{k: torch.as_tensor(v) for k, v in params.items()})
  assert isinstance(gen, smap.GenericPairNeighborListFn)
  nb_t = _wrap(nb_o, fmt)
  Rt = torch.as_tensor(R)
  np.testing.assert_allclose(float(gen(Rt, neighbor=nb_t)), E_o, rtol=1e-10)
  F_t = quantity.force(gen)(Rt, neighbor=nb_t).numpy()
  np.testing.assert_allclose(F_t, F_o, rtol=1e-8, atol=1e-9 * np.abs(F_o).max())
  if fmt != 'OrderedSparse':
    Ea_o = oenergy.pair_neighbor_list_energy(pot, d_o, R, nb_o, species=sp, per_particle=True, **params)
    gen1 = smap.pair_neighbor_list(fn, d_t, species=torch.as_tensor(sp), reduce_axis=(1,),
                                   **{k: torch.as_tensor(v) for k, v in params.items()})
    np.testing.assert_allclose(gen1(Rt, neighbor=nb_t).numpy(), Ea_o, rtol=1e-9, atol=1e-12)
  else:
    gen1 = smap.pair_neighbor_list(fn, d_t, reduce_axis=(1,))
    with pytest.raises(ValueError):
      gen1(Rt, neighbor=nb_t)


def test_pressure_and_stress_goldens_through_autograd():
  """quantity.pressure / stress (quantity.py:202-282) for a generic energy function:
  jammed soft-sphere pressure 0.06307342050945483 and the LAMMPS LJ stress tensor."""
  from jax_md_b200 import energy, quantity, smap, space
  s = np.load(os.path.join(util.GOLDEN, 'jammed_state.npz'))
  R = torch.as_tensor(s['real_position'])
  L = float(s['box'][0, 0])
  d_o, _ = ospace.periodic(L)
  d_t, _ = space.periodic(L)
  nb_o = opart.neighbor_list(d_o, L, np.float64(np.max(s['sigma'])), np.float64(0.0),
                             format=opart.Dense).allocate(s['real_position'])
  gen = smap.pair_neighbor_list(lambda dr, sigma=1.0, **kw: energy.soft_sphere(dr, sigma), d_t,
                                species=torch.as_tensor(s['species']), sigma=torch.as_tensor(s['sigma']))
  nb_t = _wrap(nb_o, 'Dense')
  np.testing.assert_allclose(float(gen(R, neighbor=nb_t)), G['jammed_energy'], rtol=1e-10)
  P = quantity.pressure(gen, R, L, neighbor=nb_t)
  np.testing.assert_allclose(float(P), G['jammed_pressure'], rtol=1e-9)

  s = np.load(os.path.join(util.GOLDEN, 'lammps_lj.npz'))
  box = np.float32(s['box'])
  Rn = (s['R'] * box).astype(np.float64)
  r = s['stress_row']
  C = np.array([[r[0], r[3], r[4]], [r[3], r[1], r[5]], [r[4], r[5], r[2]]])
  d_o, _ = ospace.periodic(box)
  d_t, _ = space.periodic(box)
  nb_o = opart.neighbor_list(d_o, box, np.float32(2.5), np.float32(0.0), format=opart.Dense).allocate(Rn)
  gen = smap.pair_neighbor_list(lambda dr, **kw: energy.lennard_jones(dr), d_t)
  S = quantity.stress(gen, torch.as_tensor(Rn), box, velocity=torch.as_tensor(s['V'].astype(np.float64)),
                      neighbor=_wrap(nb_o, 'Dense')).numpy()
  np.testing.assert_allclose(S, C, rtol=5e-5, atol=5e-5)
  assert quantity.volume(3, 2.0) == 8.0
  assert float(quantity.volume(2, torch.tensor([2.0, 3.0]))) == 6.0
  assert abs(float(quantity.volume(2, torch.tensor([[2.0, 1.0], [0.0, 3.0]]))) - 6.0) < 1e-12


def test_neighbor_list_dataclass_contract():
  """partition.py:684-737 consumers: dataclasses.replace, .set, frozen, idx is a
  plain attribute read when there is no device workspace."""
  from jax_md_b200 import partition as P
  nb = P.NeighborList(torch.arange(6).reshape(2, 3), torch.zeros(2, 3),
                      P.PartitionError(torch.tensor(3, dtype=torch.uint8)), None, 3, P.Dense,
                      None, None, None, None)
  assert nb.idx.shape == (2, 3)
  nb2 = dataclasses.replace(nb, max_occupancy=7)
  assert nb2.max_occupancy == 7 and nb2.idx is nb.idx
  assert nb.set(idx='x').idx == 'x'
  with pytest.raises(dataclasses.FrozenInstanceError):
    nb.idx = 1
  assert int(nb.did_buffer_overflow) == 3 and int(nb.cell_size_too_small) == 0
  assert not nb.internal_list_is_current
  assert str(nb.error) == 'Partition Error: Neighbor list buffer overflow.'


@pytest.mark.parametrize('fmt', FORMATS)
def test_generic_path_asymmetric_species_table(fmt):
  """smap.py:794-797 lookup orientation (Dense p[s_row, s_neigh]; sparse
  p[s_idx0, s_idx1]) told apart by an asymmetric table; forces are the true
  gradient (both rows of a pair contribute)."""
  from jax_md_b200 import energy, quantity, smap, space
  R, L = util.fcc(5, dtype=np.float64)
  R = util.jitter(R, L, 0.06)
  sp = (np.arange(len(R)) % 2).astype(np.int32)
  sig = np.array([[1.0, 0.9], [1.15, 1.05]])
  d_o, _ = ospace.periodic(L)
  d_t, _ = space.periodic(L)
  nb_o = opart.neighbor_list(d_o, L, np.float32(1.4), np.float32(0.2), format=opart.Format[fmt]).allocate(R)
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(
      oenergy.PairPotential('soft_sphere'), d_o, R, nb_o, species=sp, want_grads=True,
      sigma=sig, epsilon=np.float64(1.0), alpha=np.float64(2.0))
  gen = smap.pair_neighbor_list(lambda dr, sigma=1.0, **kw: energy.soft_sphere(dr, sigma), d_t,
                                species=torch.as_tensor(sp), sigma=torch.as_tensor(sig))
  Rt = torch.as_tensor(R)
  nb_t = _wrap(nb_o, fmt)
  np.testing.assert_allclose(float(gen(Rt, neighbor=nb_t)), E_o, rtol=1e-12)
  np.testing.assert_allclose(quantity.force(gen)(Rt, neighbor=nb_t).numpy(), F_o,
                             rtol=1e-9, atol=1e-12 * np.abs(F_o).max())
  # and the transposed table gives a different energy (the test can tell them apart)
  gen_T = smap.pair_neighbor_list(lambda dr, sigma=1.0, **kw: energy.soft_sphere(dr, sigma), d_t,
                                  species=torch.as_tensor(sp), sigma=torch.as_tensor(sig.T.copy()))
  if fmt == 'OrderedSparse':
    assert abs(float(gen_T(Rt, neighbor=nb_t)) - E_o) > 1e-6 * abs(E_o)
