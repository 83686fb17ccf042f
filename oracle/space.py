"""Oracle: spaces (reference `jax_md/space.py`).  Test infrastructure only.

NumPy >= 2 (NEP 50) scalar promotion matches JAX's weak-type promotion for the
Python-scalar cases used here, so the expressions are kept literally as the
reference writes them.
"""
import numpy as np

f32 = np.float32
f64 = np.float64


def pairwise_displacement(Ra, Rb):
  """space.py:189-210 -- plain difference."""
  return Ra - Rb


def periodic_displacement(side, dR):
  """space.py:213-224 -- `mod(dR + side/2, side) - side/2` with the reference's
  exact operand order (`side * f32(0.5)` first, `f32(0.5) * side` second)."""
  return np.mod(dR + side * f32(0.5), side) - f32(0.5) * side


def square_distance(dR):
  """space.py:227-235 -- sum of squares over the last axis.  For a length-2/3
  axis NumPy reduces sequentially: (x*x + y*y) + z*z, each op rounded."""
  sq = dR ** 2
  out = sq[..., 0]
  for k in range(1, dR.shape[-1]):
    out = out + sq[..., k]
  return out


def distance(dR):
  """space.py:238-247 -- sqrt with 0 -> 0 (safe_mask, util.py:85-88)."""
  dr = square_distance(dR)
  mask = dr > 0
  return np.where(mask, np.sqrt(np.where(mask, dr, 0)), 0).astype(dr.dtype)


def periodic_shift(side, R, dR):
  """space.py:250-252."""
  return np.mod(R + dR, side)


def raw_transform(box, R):
  """space.py:128-152."""
  box = np.asarray(box)
  if box.size == 1:
    return R * box
  if box.ndim == 1:
    return R * box
  # 'ij,...j->...i'.  XLA does not pin the summation order of this contraction (parity for matrix
  # boxes is unpinned beyond rounding); the restatement and the kernels both sum j = 0, 1, 2 with
  # separately rounded products, which makes them comparable bit for bit.
  R = np.asarray(R)
  out = np.empty(R.shape, np.result_type(box, R))
  for i in range(box.shape[0]):
    acc = box[i, 0] * R[..., 0]
    for j in range(1, box.shape[1]):
      acc = acc + box[i, j] * R[..., j]
    out[..., i] = acc
  return out


def free():
  """space.py:258-272."""
  def displacement_fn(Ra, Rb, perturbation=None, **unused):
    dR = pairwise_displacement(Ra, Rb)
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **unused):
    return R + dR
  return displacement_fn, shift_fn


def periodic(side, wrapped=True):
  """space.py:275-329."""
  def displacement_fn(Ra, Rb, perturbation=None, **unused):
    if 'box' in unused:
      raise ValueError('`space.periodic` does not accept a box argument.')
    dR = periodic_displacement(side, pairwise_displacement(Ra, Rb))
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **unused):
    if wrapped:
      return periodic_shift(side, R, dR)
    return R + dR
  return displacement_fn, shift_fn


def inverse(box):
  """space.py:110-121."""
  box = np.asarray(box)
  if box.size == 1 or box.ndim == 1:
    return 1 / box
  return np.linalg.inv(box)


def periodic_general(box, fractional_coordinates=True, wrapped=True):
  """space.py:332-472.  (Matrix boxes go through einsum, whose summation order XLA does
  not pin: parity for triclinic boxes is unpinned; diagonal boxes are exact.)"""
  inv_box = inverse(box)

  def displacement_fn(Ra, Rb, perturbation=None, **kwargs):
    _box, _inv_box = box, inv_box
    if 'box' in kwargs:
      _box = kwargs['box']
      if not fractional_coordinates:
        _inv_box = inverse(_box)
    if 'new_box' in kwargs:
      _box = kwargs['new_box']
    if not fractional_coordinates:
      Ra = raw_transform(_inv_box, Ra)
      Rb = raw_transform(_inv_box, Rb)
    dR = periodic_displacement(f32(1.0), pairwise_displacement(Ra, Rb))
    dR = raw_transform(_box, dR)
    if perturbation is not None:
      dR = raw_transform(perturbation, dR)
    return dR

  def shift_fn(R, dR, **kwargs):
    if not fractional_coordinates and not wrapped:
      return R + dR
    _box, _inv_box = box, inv_box
    if 'box' in kwargs:
      _box = kwargs['box']
      _inv_box = inverse(_box)
    if 'new_box' in kwargs:
      _box = kwargs['new_box']
    dR = raw_transform(_inv_box, dR)
    if not fractional_coordinates:
      R = raw_transform(_inv_box, R)
    R = periodic_shift(f32(1.0), R, dR) if wrapped else R + dR
    if not fractional_coordinates:
      R = raw_transform(_box, R)
    return R
  return displacement_fn, shift_fn


def metric_sq(displacement_fn):
  """partition.py:564-587 for a displacement function."""
  return lambda Ra, Rb, **kw: square_distance(displacement_fn(Ra, Rb, **kw))
