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
  general: bool = False     # space.periodic_general (orthorhombic box)
  fractional: bool = False  # ... positions stored in the unit cube


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
    side = _box_diagonal(spec.side) if spec.general else spec.side
    if spec.general and isinstance(side, np.ndarray) and side.ndim == 2:
      return _triclinic_struct(st, spec, side, dim, np_dtype)
    full, half = _side_vectors(side, dim, np_dtype)
    for k in range(dim):
      st.side[k] = float(full[k])
      st.half[k] = float(half[k])
    if spec.general:
      st.general = 1
      st.fractional = 1 if spec.fractional else 0
      # space.inverse: 1 / box, rounded in the BOX's precision when it is a typed array (an f32 box
      # next to f64 positions keeps its f32 inverse in the reference), else in the run's
      sd = np.asarray(side)
      inv_dtype = sd.dtype.type if sd.dtype in (np.float32, np.float64) and not isinstance(side, float) else np_dtype
      inv = (inv_dtype(1) / np.broadcast_to(sd.astype(inv_dtype), (dim,))).astype(np.float64)
      for k in range(dim):
        st.inv_box[k] = float(inv[k])
  return st


def _box_diagonal(box):
  """Scalar / vector / DIAGONAL-matrix box -> scalar or vector (the orthorhombic kernels);
  a matrix with off-diagonal elements is returned as it is (triclinic kernels)."""
  if isinstance(box, torch.Tensor):
    box = box.detach().cpu().numpy()
  b = np.asarray(box) if not isinstance(box, (float, int)) else box
  if isinstance(b, np.ndarray) and b.ndim == 2:
    if not np.array_equal(np.diag(np.diag(b)), b):
      return b
    return np.diag(b).copy()
  return b


def is_triclinic(box) -> bool:
  b = _box_diagonal(box)
  return isinstance(b, np.ndarray) and b.ndim == 2


def _triclinic_struct(st, spec, box, dim, np_dtype):
  """jmd_space_t for a full-matrix box: real = box @ fractional (space.py:128-150).  The inverse
  is formed in the box's own precision like `space.inverse` (space.py:110-121)."""
  if box.shape != (dim, dim):
    raise ValueError(f'box matrix must be [{dim}, {dim}], found {box.shape}')
  mdt = box.dtype.type if box.dtype in (np.float32, np.float64) else np_dtype
  H = box.astype(mdt)
  Hi = np.linalg.inv(H).astype(mdt)
  st.general, st.triclinic = 1, 1
  st.fractional = 1 if spec.fractional else 0
  for i in range(dim):
    for j in range(dim):
      st.box_m[3 * i + j] = float(np_dtype(H[i, j]))
      st.inv_box_m[3 * i + j] = float(np_dtype(Hi[i, j]))
  for k in range(dim):
    st.side[k] = float(H[k, k])
    st.inv_box[k] = float(1.0 / H[k, k])
    # |d_j| <= half[j] for all j  =>  |(H^-1 d)_i| <= 1/2: the raw difference is the minimum image
    col = np.abs(Hi[:, k].astype(np.float64)).max()
    st.half[k] = 0.5 / (dim * col) * (1.0 - 1e-3)
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


def inverse(box):
  """space.py:110-121."""
  b = box if isinstance(box, torch.Tensor) else torch.as_tensor(np.asarray(box))
  if b.numel() == 1 or b.ndim == 1:
    return 1 / b
  if b.ndim == 2:
    return torch.linalg.inv(b)
  raise ValueError(f'Box must be either: a scalar, a vector, or a matrix. Found {box}.')


def transform(box, R):
  """space.py:156-186 (forward value; gradients w.r.t. R pass through unscaled like the
  reference's custom JVP)."""
  return raw_transform(box, R)


def periodic_general(box, fractional_coordinates=True, wrapped=True):
  """space.py:332-472: periodic boundary conditions on `box * [0, 1]^d`.  The returned
  closures evaluate the reference's formulas with torch ops (host glue, tests); the tag
  lets the kernels inline the same arithmetic.  Kernel support: scalar, vector and
  diagonal-matrix boxes, `box=` overrides as host values."""
  def _b(x, like):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device=like.device, dtype=like.dtype if t.is_floating_point() else None)

  def displacement_fn(Ra, Rb, perturbation=None, **kwargs):
    _box = kwargs.get('new_box', kwargs.get('box', box))
    _inv_src = kwargs.get('box', box)
    bt = _b(_box, Ra)
    if not fractional_coordinates:
      it = inverse(_b(_inv_src, Ra))
      Ra, Rb = raw_transform(it, Ra), raw_transform(it, Rb)
    dR = torch.remainder(Ra - Rb + 0.5, 1.0) - 0.5                 # periodic_displacement(f32(1.0), .)
    dR = raw_transform(bt, dR)
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **kwargs):
    if not fractional_coordinates and not wrapped:
      return R + dR
    _box = kwargs.get('new_box', kwargs.get('box', box))
    bt = _b(_box, R)
    it = inverse(_b(kwargs.get('box', box), R))
    dR = raw_transform(it, dR)
    if not fractional_coordinates:
      R = raw_transform(it, R)
    R = torch.remainder(R + dR, 1.0) if wrapped else R + dR
    if not fractional_coordinates:
      R = raw_transform(bt, R)
    return R
  spec = SpaceSpec(_lib.SPACE_PERIODIC, box, wrapped, True, bool(fractional_coordinates))
  displacement_fn._jmd_space = spec
  shift_fn._jmd_space = spec
  return displacement_fn, shift_fn


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
