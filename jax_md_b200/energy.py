"""Energies on the B200 hot path: drop-in for the neighbour-list factories of
the reference `jax_md/energy.py`:

  soft_sphere_neighbor_list        energy.py:200-243
  lennard_jones_neighbor_list      energy.py:300-343
  morse_neighbor_list              energy.py:400-446
  stillinger_weber_neighbor_list   energy.py:962-1014

Each returns `(neighbor_fn, energy_fn)` and accepts the reference's plug-in
kwargs `neighbor_list_fn=` / `pair_neighbor_list_fn=` (defaulting to this
package's CUDA-backed ones).  The elementwise functional forms are also
exported (torch ops, host-side glue only) and carry the tag the fused kernels
dispatch on.
"""
import ctypes as C
from functools import wraps

import numpy as np
import torch

from . import _lib, partition, smap, space
from .partition import NeighborListFormat
from .util import maybe_downcast, np_max

f32 = np.float32


# ----------------------------------------------------------------------------
# functional forms
# ----------------------------------------------------------------------------

def soft_sphere(dr, sigma=1, epsilon=1, alpha=2, **unused_kwargs):
  """energy.py:125-173."""
  dr = dr / sigma
  U = epsilon / alpha * torch.clamp(1.0 - dr, min=0) ** alpha
  return torch.where(dr < 1.0, U, torch.zeros_like(U))


soft_sphere._jmd_potential = dict(kind=_lib.POT_SOFT_SPHERE, r_onset=None,
                                  r_cutoff=None)


def lennard_jones(dr, sigma=1, epsilon=1, **unused_kwargs):
  """energy.py:246-272."""
  idr = sigma / dr
  idr = idr * idr
  idr6 = idr * idr * idr
  idr12 = idr6 * idr6
  return torch.nan_to_num(4 * epsilon * (idr12 - idr6))


lennard_jones._jmd_potential = dict(kind=_lib.POT_LJ, r_onset=None, r_cutoff=None)


def morse(dr, sigma=1.0, epsilon=5.0, alpha=5.0, **unused_kwargs):
  """energy.py:346-371."""
  U = epsilon * (1 - torch.exp(-alpha * (dr - sigma))) ** 2 - epsilon
  return torch.nan_to_num(U)


morse._jmd_potential = dict(kind=_lib.POT_MORSE, r_onset=None, r_cutoff=None)


def multiplicative_isotropic_cutoff(fn, r_onset, r_cutoff):
  """energy.py:534-580."""
  r_c = float(r_cutoff ** f32(2))
  r_o = float(r_onset ** f32(2))

  def smooth_fn(dr):
    r = dr ** 2
    inner = torch.where(dr < r_cutoff,
                        (r_c - r) ** 2 * (r_c + 2 * r - 3 * r_o) / (r_c - r_o) ** 3,
                        torch.zeros_like(dr))
    return torch.where(dr < r_onset, torch.ones_like(dr), inner)

  @wraps(fn)
  def cutoff_fn(dr, *args, **kwargs):
    return smooth_fn(dr) * fn(dr, *args, **kwargs)
  pot = getattr(fn, '_jmd_potential', None)
  if pot is not None:
    if pot.get('r_cutoff') is not None:
      raise NotImplementedError('nested cutoffs')
    cutoff_fn._jmd_potential = dict(pot, r_onset=r_onset, r_cutoff=r_cutoff)
  else:
    cutoff_fn.__dict__.pop('_jmd_potential', None)
  return cutoff_fn


# ----------------------------------------------------------------------------
# neighbour-list factories
# ----------------------------------------------------------------------------

def soft_sphere_neighbor_list(displacement_or_metric, box_size, species=None,
                              sigma=1.0, epsilon=1.0, alpha=2.0,
                              dr_threshold=0.2, per_particle=False,
                              fractional_coordinates=False,
                              format=partition.OrderedSparse,
                              neighbor_list_fn=partition.neighbor_list,
                              pair_neighbor_list_fn=smap.pair_neighbor_list,
                              **neighbor_kwargs):
  """energy.py:200-243."""
  sigma = maybe_downcast(sigma)
  epsilon = maybe_downcast(epsilon)
  alpha = maybe_downcast(alpha)
  list_cutoff = np_max(sigma)
  dr_threshold = maybe_downcast(dr_threshold)
  neighbor_fn = neighbor_list_fn(
      displacement_or_metric, box_size, list_cutoff, dr_threshold,
      fractional_coordinates=fractional_coordinates, format=format,
      **neighbor_kwargs)
  energy_fn = pair_neighbor_list_fn(
      soft_sphere,
      space.canonicalize_displacement_or_metric(displacement_or_metric),
      ignore_unused_parameters=True, species=species, sigma=sigma,
      epsilon=epsilon, alpha=alpha,
      reduce_axis=(1,) if per_particle else None,
      fractional_coordinates=fractional_coordinates)
  return neighbor_fn, energy_fn


def lennard_jones_neighbor_list(displacement_or_metric, box_size, species=None,
                                sigma=1.0, epsilon=1.0, r_onset=2.0,
                                r_cutoff=2.5, dr_threshold=0.5,
                                per_particle=False,
                                fractional_coordinates=False,
                                format=partition.OrderedSparse,
                                neighbor_list_fn=partition.neighbor_list,
                                pair_neighbor_list_fn=smap.pair_neighbor_list,
                                **neighbor_kwargs):
  """energy.py:300-343."""
  sigma = maybe_downcast(sigma)
  epsilon = maybe_downcast(epsilon)
  r_onset = maybe_downcast(r_onset) * np_max(sigma)
  r_cutoff = maybe_downcast(r_cutoff) * np_max(sigma)
  dr_threshold = maybe_downcast(dr_threshold)
  neighbor_fn = neighbor_list_fn(
      displacement_or_metric, box_size, r_cutoff, dr_threshold,
      fractional_coordinates=fractional_coordinates, format=format,
      **neighbor_kwargs)
  energy_fn = pair_neighbor_list_fn(
      multiplicative_isotropic_cutoff(lennard_jones, r_onset, r_cutoff),
      space.canonicalize_displacement_or_metric(displacement_or_metric),
      ignore_unused_parameters=True, species=species, sigma=sigma,
      epsilon=epsilon, reduce_axis=(1,) if per_particle else None,
      fractional_coordinates=fractional_coordinates)
  return neighbor_fn, energy_fn


def morse_neighbor_list(displacement_or_metric, box_size, species=None,
                        sigma=1.0, epsilon=5.0, alpha=5.0, r_onset=2.0,
                        r_cutoff=2.5, dr_threshold=0.5, per_particle=False,
                        fractional_coordinates=False,
                        format=partition.OrderedSparse,
                        neighbor_list_fn=partition.neighbor_list,
                        pair_neighbor_list_fn=smap.pair_neighbor_list,
                        **neighbor_kwargs):
  """energy.py:400-446 (r_onset / r_cutoff are NOT scaled by sigma, :421-422)."""
  sigma = maybe_downcast(sigma)
  epsilon = maybe_downcast(epsilon)
  alpha = maybe_downcast(alpha)
  r_onset = maybe_downcast(r_onset)
  r_cutoff = maybe_downcast(r_cutoff)
  dr_threshold = maybe_downcast(dr_threshold)
  neighbor_fn = neighbor_list_fn(
      displacement_or_metric, box_size, r_cutoff, dr_threshold,
      fractional_coordinates=fractional_coordinates, format=format,
      **neighbor_kwargs)
  energy_fn = pair_neighbor_list_fn(
      multiplicative_isotropic_cutoff(morse, r_onset, r_cutoff),
      space.canonicalize_displacement_or_metric(displacement_or_metric),
      ignore_unused_parameters=True, species=species, sigma=sigma,
      epsilon=epsilon, alpha=alpha,
      reduce_axis=(1,) if per_particle else None,
      fractional_coordinates=fractional_coordinates)
  return neighbor_fn, energy_fn


class StillingerWeberFn:
  """energy_fn of `stillinger_weber_neighbor_list` (energy.py:994-1012) backed
  by csrc/jmd_sw.cu; also serves -dE/dR (quantity.force) from the same kernel."""
  _jmd_fused = 'sw'

  def __init__(self, displacement, params):
    self.spec = space.get_spec(displacement)
    self.params = params

  def _struct(self):
    sw = _lib.SwT()
    for k, v in self.params.items():
      setattr(sw, k, float(v))
    return sw

  def launch(self, R, neighbor, momentum=None, mass=None, dt_2=0.0,
             dt_dev=None, red=None, refresh_positions=True, **unused):
    if neighbor.format is not partition.Dense:
      raise NotImplementedError('Stillinger-Weber potential only implemented '
                                'with Dense neighbor lists.')
    ws = neighbor._ws
    if ws is None or not neighbor.internal_list_is_current:
      raise NotImplementedError('foreign NeighborList')
    if ws.dim != 3:
      raise ValueError('Stillinger-Weber needs 3-d positions')
    ws.set_species(None)
    if refresh_positions:
      _lib.call('jmd_nbr_pack', ws.ref(), _lib.ptr(R.contiguous()), _lib.stream())
    force = torch.empty_like(R)
    if red is None:
      red = torch.zeros(_lib.RED_COUNT, dtype=torch.float64, device=R.device)
    partials = smap.Scratch.get(ws.n, R.device)
    sw = self._struct()
    mass_is_array = 1 if (mass is not None and mass.numel() > 1) else 0
    scratch = ws.t.get('sw_scratch')
    need = (ws.c.m_int + 1) * ws.c.n_pad
    if scratch is None or scratch.numel() < need:
      scratch = ws.buf_plain('sw_scratch', (need,), torch.int32)
    _lib.call('jmd_sw_force', ws.ref(), C.byref(sw), _lib.ptr(scratch), _lib.ptr(force),
              _lib.ptr(red), _lib.ptr(partials), _lib.ptr(momentum),
              _lib.ptr(mass), mass_is_array, float(dt_2), _lib.ptr(dt_dev),
              _lib.stream())
    return dict(force=force, red=red)

  def force(self, R, neighbor=None, **kwargs):
    return self.launch(R, neighbor)['force']

  def __call__(self, R, neighbor=None, **kwargs):
    if neighbor is None:
      raise TypeError('energy_fn(R, neighbor=...) needs a NeighborList')
    fn = self
    if torch.is_grad_enabled() and R.requires_grad:
      class _E(torch.autograd.Function):
        @staticmethod
        def forward(ctx, Rin):
          out = fn.launch(Rin.detach(), neighbor)
          ctx.force = out['force']
          return out['red'][_lib.RED_ENERGY].to(Rin.dtype)

        @staticmethod
        def backward(ctx, g):
          return -(g * ctx.force)
      return _E.apply(R)
    out = self.launch(R, neighbor)
    return out['red'][_lib.RED_ENERGY].to(R.dtype)


def stillinger_weber_neighbor_list(displacement, box_size, sigma=2.0951,
                                   A=7.049556277, B=0.6022245584, lam=21.0,
                                   gamma=1.2, epsilon=2.16826,
                                   three_body_strength=1.0, cutoff=3.77118,
                                   dr_threshold=0.5,
                                   fractional_coordinates=False,
                                   format=partition.Dense,
                                   neighbor_list_fn=partition.neighbor_list,
                                   **neighbor_kwargs):
  """energy.py:962-1014 (`fractional_coordinates` is accepted and not
  forwarded, like the reference :985-992)."""
  neighbor_fn = neighbor_list_fn(displacement, box_size, cutoff, dr_threshold,
                                 format=format, **neighbor_kwargs)
  params = dict(sigma=sigma, A=A, B=B, lam=lam, gamma=gamma, epsilon=epsilon,
                three_body_strength=three_body_strength, cutoff=cutoff)
  return neighbor_fn, StillingerWeberFn(displacement, params)
