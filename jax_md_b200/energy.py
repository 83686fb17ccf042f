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


def dsf_coulomb(r, Q_sq, alpha=0.25, cutoff=8.0):
  """energy.py:578-597: damped-shifted-force Coulomb."""
  qqr2e = 332.06371
  import math
  cutoffsq = cutoff * cutoff
  erfcc = math.erfc(alpha * cutoff)
  erfcd = math.exp(-alpha * alpha * cutoffsq)
  f_shift = -(erfcc / cutoffsq + 2 / math.sqrt(math.pi) * alpha * erfcd / cutoff)
  e_shift = erfcc / cutoff - f_shift * cutoff
  e = qqr2e * Q_sq / r * (torch.erfc(alpha * r) - r * e_shift - r ** 2 * f_shift)
  return torch.where(r < cutoff, e, torch.zeros_like(e))


def bks(dr, Q_sq, exp_coeff, exp_decay, attractive_coeff, repulsive_coeff, coulomb_alpha, cutoff,
        **unused_kwargs):
  """energy.py:600-650: Beest-Kramer-van Santen silica potential (Buckingham form + DSF
  Coulomb + r^-24 repulsion)."""
  safe = torch.where(dr > 0, dr, torch.ones_like(dr))
  e = (dsf_coulomb(safe, Q_sq, coulomb_alpha, cutoff) + exp_coeff * torch.exp(-safe / exp_decay)
       + attractive_coeff / safe ** 6 + repulsive_coeff / safe ** 24)
  return torch.where((dr < cutoff) & (dr > 0), e, torch.zeros_like(e))


CHARGE_OXYGEN = -0.977476019
CHARGE_SILICON = 1.954952037
BKS_SILICA_DICT = {
    'Q_sq': [[CHARGE_SILICON ** 2, CHARGE_SILICON * CHARGE_OXYGEN],
             [CHARGE_SILICON * CHARGE_OXYGEN, CHARGE_OXYGEN ** 2]],
    'exp_coeff': [[0, 471671.1243], [471671.1243, 23138.64826]],
    'exp_decay': [[1, 0.19173537], [0.19173537, 0.356855265]],
    'attractive_coeff': [[0, -2156.074422], [-2156.074422, -1879.223108]],
    'repulsive_coeff': [[78940848.06, 668.7557239], [668.7557239, 2605.841269]],
    'coulomb_alpha': 0.25,
}


def bks_neighbor_list(displacement_or_metric, box_size, species, Q_sq, exp_coeff, exp_decay,
                      attractive_coeff, repulsive_coeff, coulomb_alpha, cutoff, dr_threshold=0.8,
                      fractional_coordinates=False, format=partition.OrderedSparse,
                      neighbor_list_fn=partition.neighbor_list,
                      pair_neighbor_list_fn=smap.pair_neighbor_list, **neighbor_kwargs):
  """energy.py:686-735.  The neighbour list is the CUDA one; the pair sum takes the
  generic (torch-composed) `smap.pair_neighbor_list` path: BKS has no fused kernel."""
  def table(x):
    return torch.as_tensor(np.asarray(maybe_downcast(np.asarray(x, np.float64))))
  neighbor_fn = neighbor_list_fn(displacement_or_metric, box_size, cutoff, maybe_downcast(dr_threshold),
                                 fractional_coordinates=fractional_coordinates, format=format,
                                 **neighbor_kwargs)
  energy_fn = pair_neighbor_list_fn(
      bks, space.canonicalize_displacement_or_metric(displacement_or_metric), species=species,
      ignore_unused_parameters=True, Q_sq=table(Q_sq), exp_coeff=table(exp_coeff),
      exp_decay=table(exp_decay), attractive_coeff=table(attractive_coeff),
      repulsive_coeff=table(repulsive_coeff), coulomb_alpha=coulomb_alpha, cutoff=cutoff,
      fractional_coordinates=fractional_coordinates)
  return neighbor_fn, energy_fn


def _bks_silica_self(Q_sq, alpha, cutoff):
  """energy.py:760-771."""
  import math
  cutoffsq = cutoff * cutoff
  erfcc = math.erfc(alpha * cutoff)
  erfcd = math.exp(-alpha * alpha * cutoffsq)
  f_shift = -(erfcc / cutoffsq + 2.0 / math.sqrt(math.pi) * alpha * erfcd / cutoff)
  e_shift = erfcc / cutoff - f_shift * cutoff
  return -(e_shift / 2.0 + alpha / math.sqrt(math.pi)) * Q_sq * 332.06371


def bks_silica_neighbor_list(displacement_or_metric, box_size, species, cutoff=8.0,
                             fractional_coordinates=False, format=partition.OrderedSparse,
                             **neighbor_kwargs):
  """energy.py:800-835: BKS for SiO2 incl. the self-energy terms."""
  neighbor_fn, pair_fn = bks_neighbor_list(displacement_or_metric, box_size, species, cutoff=cutoff,
                                           fractional_coordinates=fractional_coordinates, format=format,
                                           **BKS_SILICA_DICT, **neighbor_kwargs)
  sp = torch.as_tensor(species)
  N_0, N_1 = int((sp == 0).sum()), int((sp == 1).sum())
  e_self = N_0 * _bks_silica_self(CHARGE_SILICON ** 2, 0.25, cutoff) + \
      N_1 * _bks_silica_self(CHARGE_OXYGEN ** 2, 0.25, cutoff)

  def energy_fn(R, neighbor=None, **kwargs):
    return pair_fn(R, neighbor, **kwargs) + e_self
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
             dt_dev=None, red=None, refresh_positions=True, box=None, **unused):
    if neighbor._ws is not None:
      neighbor._ws.set_box(self.spec, box)                  # periodic_general: box override
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
    need = int(_lib.load().jmd_sw_scratch_ints(ws.ref()))
    if scratch is None or scratch.numel() < need:
      scratch = ws.buf_plain('sw_scratch', (need,), torch.int32)
    _lib.call('jmd_sw_force', ws.ref(), C.byref(sw), _lib.ptr(scratch), _lib.ptr(force),
              _lib.ptr(red), _lib.ptr(partials), _lib.ptr(momentum),
              _lib.ptr(mass), mass_is_array, float(dt_2), _lib.ptr(dt_dev),
              _lib.stream())
    return dict(force=force, red=red)

  def force(self, R, neighbor=None, **kwargs):
    return self.launch(R, neighbor, box=kwargs.get('box'))['force']

  def force_and_virial(self, R, neighbor=None, **kwargs):
    """One launch -> (force, trace of `virial()`), see PairNeighborListFn.force_and_virial."""
    out = self.launch(R, neighbor, box=kwargs.get('box'))
    return out['force'], out['red'][_lib.RED_VIRIAL:_lib.RED_VIRIAL + 3].sum().to(R.dtype)

  def virial(self, R, neighbor=None, **kwargs):
    """dU/d(eps_ab) at eps = 0 for the strain (I + eps) (see PairNeighborListFn.virial):
    accumulated by k_sw next to the forces.  -> [3, 3]."""
    v = self.launch(R, neighbor, box=kwargs.get('box'))['red'][_lib.RED_VIRIAL:_lib.RED_VIRIAL + 6].to(R.dtype)
    return torch.stack([torch.stack([v[0], v[3], v[4]]), torch.stack([v[3], v[1], v[5]]),
                        torch.stack([v[4], v[5], v[2]])])

  def __call__(self, R, neighbor=None, **kwargs):
    if neighbor is None:
      raise TypeError('energy_fn(R, neighbor=...) needs a NeighborList')
    fn = self
    box = kwargs.get('box')
    if neighbor._ws is not None:
      neighbor._ws.set_box(self.spec, box)
    if torch.is_grad_enabled() and R.requires_grad:
      class _E(torch.autograd.Function):
        @staticmethod
        def forward(ctx, Rin):
          out = fn.launch(Rin.detach(), neighbor, box=box)
          ctx.force = out['force']
          return out['red'][_lib.RED_ENERGY].to(Rin.dtype)

        @staticmethod
        def backward(ctx, g):
          return -(g * ctx.force)
      return _E.apply(R)
    out = self.launch(R, neighbor, box=box)
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
  """energy.py:962-1014.  `fractional_coordinates` IS forwarded to the neighbour list: the
  reference drops it (:985-992), which leaves its cell grid degenerate for unit-cube
  positions (every atom in the corner cells: same neighbour sets, quadratic cost)."""
  neighbor_fn = neighbor_list_fn(displacement, box_size, cutoff, dr_threshold,
                                 fractional_coordinates=fractional_coordinates,
                                 format=format, **neighbor_kwargs)
  params = dict(sigma=sigma, A=A, B=B, lam=lam, gamma=gamma, epsilon=epsilon,
                three_body_strength=three_body_strength, cutoff=cutoff)
  return neighbor_fn, StillingerWeberFn(displacement, params)
