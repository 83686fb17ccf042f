"""FIRE descent: drop-in for `minimize.fire_descent` of the reference
(jax_md/minimize.py:97-226).  Rides on the fused NVE step with a device-resident
(traced) dt; the |F|, |P| and F.P reductions come out of the force+kick kernel
and the momentum mixing / dt-alpha schedule is one more kernel."""
from typing import Any

import torch

from . import _lib, dataclasses, simulate


@dataclasses.dataclass
class FireDescentState:
  """minimize.py:97-121."""
  position: Any
  momentum: Any
  force: Any
  mass: Any
  dt: Any
  alpha: Any
  n_pos: Any
  _fire: Any = None          # [dt, alpha] device buffer the kernels read


def fire_descent(energy_or_force, shift_fn, dt_start=0.1, dt_max=0.4, n_min=5,
                 f_inc=1.1, f_dec=0.5, alpha_start=0.1, f_alpha=0.99):
  """minimize.py:124-226."""
  stepper = simulate._Stepper(energy_or_force, shift_fn, dt_start)

  def init_fn(R, mass=1.0, **kwargs):
    R = R.contiguous()
    P = torch.zeros_like(R)
    F = stepper.force(R, kwargs).contiguous()
    m = simulate._canonical_mass(mass, R)
    fire = torch.tensor([dt_start, alpha_start], dtype=R.dtype, device=R.device)
    n_pos = torch.zeros((), dtype=torch.int32, device=R.device)
    return FireDescentState(R, P, F, m, fire[0], fire[1], n_pos, fire)

  def apply_fn(state, **kwargs):
    R, P, F = stepper.step(state.position, state.momentum, state.force,
                           state.mass, kwargs, dt_dev=state._fire)
    fire_out = torch.empty_like(state._fire)
    npos_out = torch.empty_like(state.n_pos)
    _lib.call('jmd_fire_mix', _lib.dtype_code(R.dtype), P.numel(), _lib.ptr(P),
              _lib.ptr(F), _lib.ptr(stepper.red(R)), _lib.ptr(state._fire),
              _lib.ptr(fire_out), _lib.ptr(state.n_pos), _lib.ptr(npos_out),
              float(dt_max), float(n_min), float(f_inc), float(f_dec),
              float(alpha_start), float(f_alpha), _lib.stream())
    return FireDescentState(R, P, F, state.mass, fire_out[0], fire_out[1],
                            npos_out, fire_out)

  return init_fn, apply_fn
