"""CPU tests of the reference's plug-in points: the energy factories call
`neighbor_list_fn=` / `pair_neighbor_list_fn=` with the reference's arguments
(energy.py:200-243, 300-343, 400-446), and `lax.fori_loop` keeps the reference
loop semantics without a device."""
import numpy as np
import pytest
import torch


class _Rec:
  def __init__(self):
    self.calls = []

  def __call__(self, *args, **kwargs):
    self.calls.append((args, kwargs))
    return ('fn', len(self.calls))


def _factories():
  from jax_md_b200 import energy, partition, space
  d, _ = space.periodic(np.float32(20.0))
  return energy, partition, d


def test_lennard_jones_factory_arguments():
  energy, partition, d = _factories()
  nl, pl = _Rec(), _Rec()
  sigma = np.array([[1.0, 1.1], [1.1, 1.2]], np.float32)
  nf, ef = energy.lennard_jones_neighbor_list(
      d, np.float32(20.0), species=np.array([0, 1]), sigma=sigma, dr_threshold=0.3,
      neighbor_list_fn=nl, pair_neighbor_list_fn=pl, capacity_multiplier=1.5)
  (args, kw), = nl.calls
  assert args[0] is d and float(args[1]) == 20.0
  np.testing.assert_allclose(float(args[2]), 2.5 * 1.2, rtol=1e-6)     # r_cutoff * max(sigma), :329
  np.testing.assert_allclose(float(args[3]), 0.3, rtol=1e-6)
  assert kw['format'] is partition.OrderedSparse and kw['fractional_coordinates'] is False
  assert kw['capacity_multiplier'] == 1.5                               # **neighbor_kwargs pass through
  (pargs, pkw), = pl.calls
  pot = pargs[0]._jmd_potential                                         # cutoff-wrapped lennard_jones
  np.testing.assert_allclose([float(pot['r_onset']), float(pot['r_cutoff'])], [2.0 * 1.2, 2.5 * 1.2], rtol=1e-6)
  assert pkw['ignore_unused_parameters'] is True and pkw['reduce_axis'] is None
  assert pkw['sigma'] is not None and 'epsilon' in pkw
  _, ef2 = energy.lennard_jones_neighbor_list(d, np.float32(20.0), per_particle=True,
                                              neighbor_list_fn=_Rec(), pair_neighbor_list_fn=pl)
  assert pl.calls[-1][1]['reduce_axis'] == (1,)


def test_soft_sphere_and_morse_factory_arguments():
  energy, partition, d = _factories()
  nl, pl = _Rec(), _Rec()
  sigma = np.array([[1.0, 1.2], [1.2, 1.4]], np.float32)
  energy.soft_sphere_neighbor_list(d, np.float32(20.0), species=np.array([0, 1]), sigma=sigma,
                                   neighbor_list_fn=nl, pair_neighbor_list_fn=pl)
  (args, kw), = nl.calls
  np.testing.assert_allclose(float(args[2]), 1.4, rtol=1e-6)            # list cutoff = max(sigma), :223
  np.testing.assert_allclose(float(args[3]), 0.2, rtol=1e-6)            # default skin
  assert kw['format'] is partition.OrderedSparse
  assert pl.calls[0][0][0] is energy.soft_sphere and pl.calls[0][1]['alpha'] == 2.0
  nl, pl = _Rec(), _Rec()
  energy.morse_neighbor_list(d, np.float32(20.0), sigma=1.3, neighbor_list_fn=nl, pair_neighbor_list_fn=pl)
  (args, kw), = nl.calls
  np.testing.assert_allclose(float(args[2]), 2.5, rtol=1e-6)            # NOT scaled by sigma, :421-422
  np.testing.assert_allclose(float(args[3]), 0.5, rtol=1e-6)
  pot = pl.calls[0][0][0]._jmd_potential
  np.testing.assert_allclose([float(pot['r_onset']), float(pot['r_cutoff'])], [2.0, 2.5], rtol=1e-6)
  assert pl.calls[0][1]['epsilon'] == 5.0 and pl.calls[0][1]['alpha'] == 5.0


def test_neighbor_list_fns_interface():
  """partition.py:740-785: iterable as (allocate, update); formats validated early;
  unsupported reference options fail loudly."""
  from jax_md_b200 import partition, space
  d, _ = space.periodic(np.float32(20.0))
  fns = partition.neighbor_list(d, np.float32(20.0), 2.5, 0.3, format=partition.Sparse)
  allocate, update = fns
  assert allocate is fns.allocate and update is fns.update
  with pytest.raises(ValueError):
    partition.neighbor_list(d, np.float32(20.0), 2.5, 0.3, format='Dense')
  with pytest.raises(ValueError):        # fractional coordinates need a periodic_general space
    partition.neighbor_list(d, np.float32(20.0), 2.5, 0.3, fractional_coordinates=True)
  dg, _ = space.periodic_general(np.float32(20.0))
  assert isinstance(partition.neighbor_list(dg, np.float32(20.0), 2.5, 0.3, fractional_coordinates=True),
                    partition.NeighborListFns)
  masked = partition.neighbor_list(d, np.float32(20.0), 2.5, 0.3, custom_mask_function=lambda idx: idx)
  assert isinstance(masked, partition.NeighborListFns)        # served by the post-mask path
  assert partition.is_sparse(partition.OrderedSparse) and not partition.is_sparse(partition.Dense)


def test_fori_loop_reference_semantics_without_device():
  """lax.fori_loop(lower, upper, body, init): eager on CPU, any carry structure."""
  from jax_md_b200 import lax
  def body(i, carry):
    x, d = carry
    return x + i, {'n': d['n'] + 1}
  x, d = lax.fori_loop(2, 7, body, (torch.zeros(3), {'n': 0}))
  assert torch.equal(x, torch.full((3,), float(2 + 3 + 4 + 5 + 6))) and d['n'] == 5
  init = (torch.ones(2), {'n': 0})
  assert lax.fori_loop(5, 5, body, init) is init


def test_dispatch_by_state_routes_on_the_position_type():
  """simulate.py:100-116."""
  from types import SimpleNamespace
  from jax_md_b200 import simulate

  class Quaternions(tuple):
    pass

  @simulate.dispatch_by_state
  def step(state, scale=1):
    return ('default', scale)

  @step.register(Quaternions)
  def _(state, scale=1):
    return ('rigid', scale)

  assert step(SimpleNamespace(position=[1.0]), scale=2) == ('default', 2)
  assert step(SimpleNamespace(position=Quaternions((1.0,))), scale=3) == ('rigid', 3)
  assert step(SimpleNamespace(position=(1.0,))) == ('default', 1)      # a plain tuple is not the subclass
