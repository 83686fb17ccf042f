"""Spaces: drop-in for the reference `jax_md/space.py` (free, periodic).

`free()` / `periodic(side)` return `(displacement_fn, shift_fn)` exactly like
space.py:258-329.  The returned closures also carry a `_jmd_space` tag so the
neighbour-list, force and integrator kernels can inline the periodic
displacement and shift (space.py:213-224, 250-252) instead of calling back into
Python.  The closures themselves are evaluated with torch ops and are only
meant for host-side glue and tests -- never for the hot path.
"""
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib

f32 = np.float32


class UnexpectedBoxException(Exception):
  pass


class SpaceSpec(NamedTuple):
  kind: int                 # _lib.SPACE_FREE / SPACE_PERIODIC
  side: Optional[object]    # as given by the user (float, np scalar, array)
  wrapped: bool


def _side_vectors(side, dim, np_dtype):
  """`side` and `f32(0.5) * side` in the position dtype (NumPy promotion ==
  JAX weak-type promotion here), broadcast to [dim]."""
  if isinstance(side, torch.Tensor):
    side = side.detach().cpu().numpy()
  s = np.asarray(side) if not isinstance(side, (float, int)) else side
  if isinstance(s, np.ndarray) and s.ndim == 2:
    raise ValueError('space.periodic takes a scalar or vector box; '
                     'periodic_general is not part of this path (SURVEY 8f-3).')
  half = f32(0.5) * s        # python float -> f32; f64 array -> f64
  full = np.broadcast_to(np.asarray(s, np_dtype), (dim,)).astype(np.float64)
  half = np.broadcast_to(np.asarray(half).astype(np_dtype), (dim,)).astype(np.float64)
  return full, half


def space_struct(spec: SpaceSpec, dim: int, torch_dtype) -> '_lib.SpaceT':
  np_dtype = np.float32 if torch_dtype == torch.float32 else np.float64
  st = _lib.SpaceT()
  st.dim = dim
  st.kind = spec.kind
  st.wrapped = 1 if spec.wrapped else 0
  if spec.kind == _lib.SPACE_PERIODIC:
    full, half = _side_vectors(spec.side, dim, np_dtype)
    for k in range(dim):
      st.side[k] = float(full[k])
      st.half[k] = float(half[k])
  return st


def _as_like(x, ref):
  if isinstance(x, torch.Tensor):
    return x.to(device=ref.device)
  return torch.as_tensor(np.asarray(x), device=ref.device)


def raw_transform(box, R):
  """space.py:128-152."""
  box = _as_like(box, R)
  if box.numel() == 1 or box.ndim == 1:
    return R * box
  return torch.einsum('ij,...j->...i', box.to(R.dtype), R)


def pairwise_displacement(Ra, Rb):
  """space.py:189-210."""
  if Ra.ndim != 1:
    raise ValueError('Can only compute displacements between vectors.')
  if Ra.shape != Rb.shape:
    raise ValueError('Can only compute displacement between vectors of equal '
                     'dimension.')
  return Ra - Rb


def periodic_displacement(side, dR):
  """space.py:213-224 (torch remainder == jnp.mod for positive side)."""
  if isinstance(side, (float, int)):
    half = float(f32(side) * f32(0.5))     # python float * f32 -> f32
    return torch.remainder(dR + half, side) - half
  s = _as_like(side, dR)
  half = (s * 0.5)
  return torch.remainder(dR + half, s) - half


def square_distance(dR):
  return torch.sum(dR ** 2, dim=-1)


def distance(dR):
  dr = square_distance(dR)
  return torch.where(dr > 0, torch.sqrt(torch.where(dr > 0, dr, torch.ones_like(dr))),
                     torch.zeros_like(dr))


def periodic_shift(side, R, dR):
  """space.py:250-252."""
  if isinstance(side, (float, int)):
    return torch.remainder(R + dR, side)
  return torch.remainder(R + dR, _as_like(side, R))


def free():
  """space.py:258-272."""
  def displacement_fn(Ra, Rb, perturbation=None, **unused_kwargs):
    dR = Ra - Rb
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **unused_kwargs):
    return R + dR
  spec = SpaceSpec(_lib.SPACE_FREE, None, False)
  displacement_fn._jmd_space = spec
  shift_fn._jmd_space = spec
  return displacement_fn, shift_fn


def periodic(side, wrapped: bool = True):
  """space.py:275-329."""
  def displacement_fn(Ra, Rb, perturbation=None, **unused_kwargs):
    if 'box' in unused_kwargs:
      raise UnexpectedBoxException(
          '`space.periodic` does not accept a box argument. Perhaps you meant '
          'to use `space.periodic_general`?')
    dR = periodic_displacement(side, Ra - Rb)
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **unused_kwargs):
    if 'box' in unused_kwargs:
      raise UnexpectedBoxException(
          '`space.periodic` does not accept a box argument. Perhaps you meant '
          'to use `space.periodic_general`?')
    if wrapped:
      return periodic_shift(side, R, dR)
    return R + dR
  spec = SpaceSpec(_lib.SPACE_PERIODIC, side, wrapped)
  displacement_fn._jmd_space = spec
  shift_fn._jmd_space = spec
  return displacement_fn, shift_fn


def periodic_general(box, fractional_coordinates=True, wrapped=True):
  raise NotImplementedError(
      'space.periodic_general is row 3 of SURVEY.md 8(f) ("next"), not part of '
      'the B200 hot path yet.')


def metric(displacement):
  """space.py:474-477."""
  fn = lambda Ra, Rb, **kwargs: distance(displacement(Ra, Rb, **kwargs))
  if hasattr(displacement, '_jmd_space'):
    fn._jmd_space = displacement._jmd_space
  return fn


def canonicalize_displacement_or_metric(displacement_or_metric):
  """space.py:505-522; keeps the space tag."""
  spec = getattr(displacement_or_metric, '_jmd_space', None)
  if spec is None:
    raise NotImplementedError(
        'Only displacement functions created by jax_md_b200.space.free() / '
        'periodic() can be inlined into the CUDA kernels.')
  return displacement_or_metric


def get_spec(fn) -> SpaceSpec:
  spec = getattr(fn, '_jmd_space', None)
  if spec is None:
    raise NotImplementedError(
        'This function was not created by jax_md_b200.space.free()/periodic(); '
        'arbitrary Python displacement/shift functions cannot be inlined into '
        'the CUDA kernels (generic path: SURVEY.md 8(f) row 1).')
  return spec


def map_product(fn):
  return lambda Ra, Rb, **kw: fn(Ra[:, None, :], Rb[None, :, :], **kw)


def map_bond(fn):
  return lambda Ra, Rb, **kw: fn(Ra, Rb, **kw)


def map_neighbor(fn):
  """space.py:494-502: evaluates fn(R_neigh, R_i)."""
  return lambda Ra, Rb, **kw: fn(Rb, Ra[:, None, :], **kw)
