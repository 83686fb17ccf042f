"""`quantity`: force / kinetic energy / temperature (jax_md/quantity.py:58-199)."""
import torch

from . import util


def force(energy_fn):
  """quantity.py:58-60: force = -grad(energy).  Fused energy functions answer
  from their force kernel directly; anything else goes through autograd."""
  fused = getattr(energy_fn, 'force', None)
  if getattr(energy_fn, '_jmd_fused', None) and fused is not None:
    return fused

  def force_fn(R, *args, **kwargs):
    Rg = R.detach().requires_grad_(True)
    with torch.enable_grad():
      E = energy_fn(Rg, *args, **kwargs)
      (g,) = torch.autograd.grad(E, Rg)
    return -g
  return force_fn


def canonicalize_force(energy_or_force_fn):
  """quantity.py:76-102: detect energy-vs-force by output shape on first use."""
  if getattr(energy_or_force_fn, '_jmd_fused', None):
    return force(energy_or_force_fn)
  _force_fn = None

  def force_fn(R, **kwargs):
    nonlocal _force_fn
    if _force_fn is None:
      with torch.no_grad():
        out = energy_or_force_fn(R, **kwargs)
      if out.ndim == 0:
        _force_fn = force(energy_or_force_fn)
      else:
        if out.shape != R.shape:
          raise ValueError('Provided function should be compatible with either '
                           'an energy or a force. Found a function whose output '
                           f'has shape {out.shape}.')
        _force_fn = energy_or_force_fn
        return out
    return _force_fn(R, **kwargs)
  return force_fn


def volume(dimension, box):
  """quantity.py:110-121."""
  if isinstance(box, (int, float)):
    return float(box) ** dimension
  b = torch.as_tensor(box) if not isinstance(box, torch.Tensor) else box
  if b.ndim == 0:
    return b ** dimension
  if b.ndim == 1:
    return torch.prod(b)
  if b.ndim == 2:
    return torch.linalg.det(b)
  raise ValueError(f'Box must be either: a scalar, a vector, or a matrix. Found {box}.')


def _dU_deps(energy_fn, position, eps0, make_perturbation, kwargs):
  """grad of U(eps) = energy_fn(R, perturbation=...) at eps = 0 by autograd
  (generic energy functions)."""
  eps = eps0.clone().requires_grad_(True)
  with torch.enable_grad():
    U = energy_fn(position, perturbation=make_perturbation(eps), **kwargs)
    (g,) = torch.autograd.grad(U, eps)
  return g


def pressure(energy_fn, position, box, kinetic_energy=0.0, **kwargs):
  """quantity.py:202-235: P = (2 KE - dU/deps) / (dim V), eps the isotropic box
  strain.  Fused neighbour-list energies answer from the virial their force
  kernel accumulates; anything else differentiates through `perturbation=`."""
  dim = position.shape[1]
  vol_0 = volume(dim, box)
  if getattr(energy_fn, '_jmd_fused', None) and not getattr(energy_fn, 'always_generic', False) \
      and getattr(kwargs.get('neighbor'), '_ws', None) is not None:
    dUdV = torch.trace(energy_fn.virial(position, box=box, **kwargs))     # U(eps) is evaluated AT `box` (:226)
  else:
    zero = torch.zeros((), dtype=position.dtype, device=position.device)
    dUdV = _dU_deps(energy_fn, position, zero, lambda e: 1 + e, kwargs)
  return 1 / (dim * vol_0) * (2 * kinetic_energy - dUdV)


def stress(energy_fn, position, box, mass=1.0, velocity=None, **kwargs):
  """quantity.py:238-282: (sum m v v^T - dU/deps) / V, eps the box strain tensor."""
  dim = position.shape[1]
  vol_0 = volume(dim, box)
  if getattr(energy_fn, '_jmd_fused', None) and not getattr(energy_fn, 'always_generic', False) \
      and getattr(kwargs.get('neighbor'), '_ws', None) is not None:
    dUdV = energy_fn.virial(position, box=box, **kwargs)
  else:
    zero = torch.zeros((dim, dim), dtype=position.dtype, device=position.device)
    eye = torch.eye(dim, dtype=position.dtype, device=position.device)
    dUdV = _dU_deps(energy_fn, position, zero, lambda e: eye + e, kwargs)
  VxV = 0.0
  if velocity is not None:
    m = _mass_like(mass, velocity)
    if isinstance(m, torch.Tensor) and m.ndim == 2:
      m = m[:, :, None]
    VxV = util.high_precision_sum(m * velocity[:, None, :] * velocity[:, :, None], axis=0)
  return 1 / vol_0 * (VxV - dUdV)


def count_dof(position):
  """quantity.py:105-108."""
  return position.numel()


def _check(unused_args, momentum, velocity):
  if unused_args:
    raise ValueError('To use the kinetic energy function, you must explicitly '
                     'pass either momentum or velocity as a keyword argument.')
  if momentum is not None and velocity is not None:
    raise ValueError('To use the kinetic energy function, you must pass either '
                     'a momentum or a velocity.')


def _mass_like(mass, q):
  if isinstance(mass, torch.Tensor):
    return mass.reshape(-1, 1) if mass.ndim == 1 and mass.numel() > 1 else mass
  return mass


def kinetic_energy(*unused_args, momentum=None, velocity=None, mass=1.0):
  """quantity.py:124-159."""
  _check(unused_args, momentum, velocity)
  q = velocity if momentum is None else momentum
  m = _mass_like(mass, q)
  k = q ** 2 * m if momentum is None else q ** 2 / m
  return 0.5 * util.high_precision_sum(k)


def temperature(*unused_args, momentum=None, velocity=None, mass=1.0):
  """quantity.py:162-199."""
  _check(unused_args, momentum, velocity)
  q = velocity if momentum is None else momentum
  m = _mass_like(mass, q)
  t = q ** 2 * m if momentum is None else q ** 2 / m
  return util.high_precision_sum(t) / count_dof(q)
