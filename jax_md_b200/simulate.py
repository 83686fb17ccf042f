"""Integrators on the B200 hot path: drop-in for `simulate.nve` and
`simulate.nvt_nose_hoover` of the reference (jax_md/simulate.py:279-317,
565-669), with the same `(init_fn, apply_fn)` contract and state field names.

When the energy function is one of this package's fused neighbour-list
energies and `shift_fn` comes from `jax_md_b200.space`, one MD step is two
kernels:
   jmd_nve_kick_drift   p += dt/2 F;  R = shift(R, dt p / m);  refresh the
                        cell-sorted float4 positions        (simulate.py:238-239)
   jmd_pair_force       F = -dE/dR over the full neighbour rows, fused with
                        p += dt/2 F and the KE/|F|^2/|P|^2/F.P reductions
                                                            (simulate.py:240-241)
Otherwise the same two integrator kernels bracket a call to the user's force
function.  No step synchronises with the host.
"""
from typing import Any

import numpy as np
import torch

from . import _lib, dataclasses, quantity, smap, space

f32 = np.float32

SUZUKI_YOSHIDA_WEIGHTS = {
    1: [1],
    3: [0.828981543588751, -0.657963087177502, 0.828981543588751],
    5: [0.2967324292201065, 0.2967324292201065, -0.186929716880426,
        0.2967324292201065, 0.2967324292201065],
    7: [0.784513610477560, 0.235573213359357, -1.17767998417887,
        1.31518632068391, -1.17767998417887, 0.235573213359357,
        0.784513610477560],
}


def _canonical_mass(mass, R):
  """simulate.canonicalize_mass (simulate.py:119-138): float stays scalar,
  [N] becomes [N, 1].  Kernels take the flat view."""
  if isinstance(mass, torch.Tensor):
    m = mass.to(device=R.device, dtype=R.dtype)
    if m.ndim == 1 and m.numel() > 1:
      m = m.reshape(-1, 1)
    elif m.ndim == 0:
      m = m.reshape(1)
    return m.contiguous()
  m = np.asarray(mass)
  if m.ndim == 0:
    return torch.full((1,), float(m), dtype=R.dtype, device=R.device)
  t = torch.as_tensor(m, dtype=R.dtype, device=R.device)
  return t.reshape(-1, 1).contiguous() if t.ndim == 1 else t.contiguous()


def _generator(key, device):
  if isinstance(key, torch.Generator):
    return key
  g = torch.Generator(device=device)
  seed = int(key) if not isinstance(key, torch.Tensor) else int(key.flatten()[0])
  g.manual_seed(seed)
  return g


def initialize_momenta(R, mass, key, kT):
  """simulate.py:141-159: Maxwell-Boltzmann, centred when N > 1."""
  g = _generator(key, R.device)
  p = torch.sqrt(mass * kT) * torch.randn(R.shape, dtype=R.dtype, device=R.device,
                                          generator=g)
  if R.shape[0] > 1:
    p = p - p.mean(dim=0, keepdim=True)
  return p


@dataclasses.dataclass
class NVEState:
  """simulate.py:249-275."""
  position: Any
  momentum: Any
  force: Any
  mass: Any

  @property
  def velocity(self):
    return self.momentum / self.mass


class _Stepper:
  """Shared machinery of nve / nvt / fire: the two-kernel velocity Verlet."""

  def __init__(self, energy_or_force_fn, shift_fn, dt):
    self.fn = energy_or_force_fn
    self.fused = getattr(energy_or_force_fn, '_jmd_fused', None)
    self.force_fn = quantity.canonicalize_force(energy_or_force_fn)
    # A shift function made by jax_md_b200.space is inlined into the drift kernel; any
    # other callable `shift_fn(R, dR, **kw)` is applied as given (simulate.py:176-188),
    # with the kick halves still on the CUDA kernels.
    self.shift_fn = shift_fn
    self.spec = getattr(shift_fn, '_jmd_space', None)
    self.generic_shift = self.spec is None
    if self.generic_shift:
      self.spec = space.SpaceSpec(_lib.SPACE_FREE, None, False)
    self.dt = float(f32(dt))
    self.dt_2 = float(f32(f32(dt) / 2))
    self._sp = {}
    self._red = {}

  def space_struct(self, R):
    key = (R.shape[1], R.dtype)
    st = self._sp.get(key)
    if st is None:
      st = space.space_struct(self.spec, R.shape[1], R.dtype)
      self._sp[key] = st
    return st

  def red(self, R):
    key = (str(R.device),)
    r = self._red.get(key)
    if r is None:
      r = torch.zeros(_lib.RED_COUNT, dtype=torch.float64, device=R.device)
      self._red[key] = r
    return r

  def force(self, R, kwargs):
    return self.force_fn(R, **kwargs)

  def step(self, R, P, F, mass, kwargs, dt=None, dt_dev=None, scale_dev=None):
    """One velocity-Verlet step (simulate.py:227-243) -> (R', P', F')."""
    import ctypes as C
    _lib.require_cuda()
    N, dim = R.shape
    dtc = _lib.dtype_code(R.dtype)
    dt_h = self.dt if dt is None else float(f32(dt))
    dt2_h = self.dt_2 if dt is None else float(f32(f32(dt) / 2))
    neighbor = kwargs.get('neighbor')
    fused = bool(self.fused) and neighbor is not None and \
        getattr(neighbor, '_ws', None) is not None and \
        neighbor.internal_list_is_current and \
        not getattr(self.fn, 'always_generic', False)
    sp = self.space_struct(R)
    if self.spec.general and kwargs.get('box') is not None:
      # periodic_general: `box=` overrides the box of the shift for this step (space.py:445-447)
      b = kwargs['box']
      b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b
      sp = space.space_struct(self.spec._replace(side=b), R.shape[1], R.dtype)
    R2 = torch.empty_like(R)
    P2 = torch.empty_like(P)
    mass_is_array = 1 if mass.numel() > 1 else 0
    red = self.red(R)
    st = _lib.stream()
    if self.generic_shift:
      return self._step_generic_shift(R, P, F, mass, kwargs, dt_h, dt2_h, dt_dev, scale_dev)
    nb_ref = None
    if fused:
      ws = neighbor._ws
      if self.fused == 'pair':
        kw = dict(kwargs)
        kw.pop('neighbor')
        _, species, params = self.fn._resolve(neighbor, kw)
        ws.set_species(self.fn._species_tensor(species, R.device))
      else:
        ws.set_species(None)
        ws.set_box(self.fn.spec, kwargs.get('box'))
      nb_ref = ws.ref()
    _lib.call('jmd_nve_kick_drift', C.byref(sp), dtc, N, nb_ref, _lib.ptr(R),
              _lib.ptr(P), _lib.ptr(F), _lib.ptr(mass), mass_is_array, dt_h,
              _lib.ptr(dt_dev), _lib.ptr(scale_dev), _lib.ptr(R2), _lib.ptr(P2),
              st)
    if fused:
      # the drift evaluated the skin predicate of R2 against this list's
      # reference positions (csrc/jmd_integrate.cu); update(R2) may reuse it
      import weakref
      ws.drift_out = (weakref.ref(R2), R2._version, ws.update_epoch)
    if fused and self.fused == 'pair':
      out = self.fn.launch(R2, neighbor, species, params, want_energy=False,
                           momentum=P2, mass=mass, dt_2=dt2_h, dt_dev=dt_dev,
                           red=red, refresh_positions=False)
      F2 = out['force']
    elif fused:
      out = self.fn.launch(R2, neighbor, momentum=P2, mass=mass, dt_2=dt2_h,
                           dt_dev=dt_dev, red=red, refresh_positions=False, box=kwargs.get('box'))
      F2 = out['force']
    else:
      F2 = self.force(R2, kwargs).contiguous()
      partials = smap.Scratch.get(N, R.device)
      _lib.call('jmd_kick_reduce', dtc, N, dim, _lib.ptr(P2), _lib.ptr(F2),
                _lib.ptr(mass), mass_is_array, dt2_h, _lib.ptr(dt_dev),
                _lib.ptr(red), _lib.ptr(partials), st)
    return R2, P2, F2


def _generic_shift_step(self, R, P, F, mass, kwargs, dt_h, dt2_h, dt_dev, scale_dev):
  """velocity Verlet around a user `shift_fn` (simulate.py:227-243): the position update is
  the user's callable, forces come from `force_fn`, second kick + reductions from the
  kick kernel."""
  dt = dt_h if dt_dev is None else dt_dev.reshape(-1)[0].to(R.dtype)
  dt_2 = dt2_h if dt_dev is None else (dt / 2)
  P1 = P * scale_dev if scale_dev is not None else P
  P1 = P1 + dt_2 * F
  space_kw = {k: v for k, v in kwargs.items() if k != 'neighbor'}
  R2 = self.shift_fn(R, dt * P1 / mass, **space_kw).contiguous()
  F2 = self.force(R2, kwargs).contiguous()
  P2 = P1.contiguous().clone()
  partials = smap.Scratch.get(R.shape[0], R.device)
  _lib.call('jmd_kick_reduce', _lib.dtype_code(R.dtype), R.shape[0], R.shape[1], _lib.ptr(P2), _lib.ptr(F2),
            _lib.ptr(mass), 1 if mass.numel() > 1 else 0, dt2_h, _lib.ptr(dt_dev),
            _lib.ptr(self.red(R)), _lib.ptr(partials), _lib.stream())
  return R2, P2, F2


_Stepper._step_generic_shift = _generic_shift_step


def dispatch_by_state(fn):
  """simulate.py:100-116: call `fn(state, ...)` unless an override was registered for the TYPE OF
  `state.position` (how the reference routes rigid-body states through the same step functions):
  `@step.register(RigidBody)`."""
  import functools
  overrides = {}

  @functools.wraps(fn)
  def call(state, *args, **kwargs):
    return overrides.get(type(state.position), fn)(state, *args, **kwargs)

  def register(oftype):
    def add(override):
      overrides[oftype] = override
    return add
  call.register = register
  return call


def nve(energy_or_force_fn, shift_fn, dt=1e-3, **sim_kwargs):
  """simulate.py:279-317."""
  stepper = _Stepper(energy_or_force_fn, shift_fn, dt)

  def init_fn(key, R, kT, mass=f32(1.0), momenta=None, **kwargs):
    R = R.contiguous()
    force = stepper.force(R, kwargs).contiguous()
    m = _canonical_mass(mass, R)
    if momenta is None:
      P = initialize_momenta(R, m, key, kT)
    else:
      P = momenta.to(dtype=R.dtype, device=R.device)
    return NVEState(R, P.contiguous(), force, m)

  def step_fn(state, **kwargs):
    _dt = kwargs.pop('dt', None)
    dt_dev = _dt if isinstance(_dt, torch.Tensor) else None
    R, P, F = stepper.step(state.position, state.momentum, state.force,
                           state.mass, kwargs,
                           dt=None if dt_dev is not None else _dt, dt_dev=dt_dev)
    return state.set(position=R, momentum=P, force=F)

  step_fn._stepper = stepper
  return init_fn, step_fn


# -- Nose-Hoover chain -----------------------------------------------------------

@dataclasses.dataclass
class NoseHooverChain:
  """simulate.py:350-374.  The chain lives in ONE device buffer laid out
  [xi | p_xi | Q | KE] (what the single-thread chain kernel reads and writes);
  `position`, `momentum`, `mass` and `kinetic_energy` are views of it.  Every
  `apply_fn` writes a NEW buffer, so earlier states keep their own chain."""
  _buf: Any
  tau: Any
  degrees_of_freedom: int = dataclasses.static_field()
  _cl: int = dataclasses.static_field(default=0)

  @property
  def position(self):
    return self._buf[0:self._cl]

  @property
  def momentum(self):
    return self._buf[self._cl:2 * self._cl]

  @property
  def mass(self):
    return self._buf[2 * self._cl:3 * self._cl]

  @property
  def kinetic_energy(self):
    return self._buf[3 * self._cl]


def _make_chain(buf, cl, tau, dof):
  return NoseHooverChain(buf, tau, dof, cl)


@dataclasses.dataclass
class NVTNoseHooverState:
  """simulate.py:538-562."""
  position: Any
  momentum: Any
  force: Any
  mass: Any
  chain: NoseHooverChain

  @property
  def velocity(self):
    return self.momentum / self.mass


def nvt_nose_hoover(energy_or_force_fn, shift_fn, dt, kT, chain_length=5,
                    chain_steps=2, sy_steps=3, tau=None, **sim_kwargs):
  """simulate.py:565-669."""
  stepper = _Stepper(energy_or_force_fn, shift_fn, dt)
  dt_f = f32(dt)
  if tau is None:
    tau = dt_f * 100
  tau = f32(tau)
  if sy_steps not in SUZUKI_YOSHIDA_WEIGHTS:
    raise ValueError('sy_steps must be 1, 3, 5 or 7')
  kT_cache = {}

  def _kT_dev(_kT, R):
    if isinstance(_kT, torch.Tensor):
      return _kT.to(device=R.device, dtype=R.dtype).reshape(1)
    key = (float(_kT), R.dtype, str(R.device))
    t = kT_cache.get(key)
    if t is None:
      t = torch.full((1,), float(_kT), dtype=R.dtype, device=R.device)
      kT_cache[key] = t
    return t

  def _half_step(buf_in, buf_out, dof, kT_t, R, ke_red):
    dtc = _lib.dtype_code(R.dtype)
    scale = torch.empty((1,), dtype=R.dtype, device=R.device)
    _lib.call('jmd_nhc_half_step', dtc, chain_length, chain_steps, sy_steps,
              float(dt_f), float(tau), int(dof),
              _lib.ptr(kT_t), _lib.ptr(buf_in), _lib.ptr(buf_out),
              None if ke_red is None else
              (ke_red.data_ptr() + 8 * _lib.RED_KINETIC),
              _lib.ptr(scale), _lib.stream())
    return scale

  def init_fn(key, R, mass=f32(1.0), momenta=None, **kwargs):
    _kT = kT if 'kT' not in kwargs else kwargs.pop('kT')
    R = R.contiguous()
    dof = quantity.count_dof(R)
    force = stepper.force(R, kwargs).contiguous()
    m = _canonical_mass(mass, R)
    if momenta is None:
      P = initialize_momenta(R, m, key, _kT)
    else:
      P = momenta.to(dtype=R.dtype, device=R.device)
    P = P.contiguous()
    KE = quantity.kinetic_energy(momentum=P, mass=m)
    buf = torch.zeros(3 * chain_length + 1, dtype=R.dtype, device=R.device)
    kT_h = float(_kT.detach().cpu()) if isinstance(_kT, torch.Tensor) else _kT
    Q = f32(kT_h) * (tau ** f32(2))                        # simulate.py:440-442
    buf[2 * chain_length:3 * chain_length] = float(Q)
    buf[2 * chain_length] = float(f32(Q * f32(dof)))
    buf[3 * chain_length] = KE
    return NVTNoseHooverState(R, P, force, m,
                              _make_chain(buf, chain_length, tau, dof))

  def apply_fn(state, **kwargs):
    _kT = kT if 'kT' not in kwargs else kwargs.pop('kT')
    R = state.position
    kT_t = _kT_dev(_kT, R)
    chain = state.chain
    dof = chain.degrees_of_freedom
    buf = torch.empty_like(chain._buf)       # the new state's chain (the old one stays intact)
    # update_mass + first half step on the carried KE (simulate.py:654-658)
    s1 = _half_step(chain._buf, buf, dof, kT_t, R, None)
    R2, P2, F2 = stepper.step(R, state.momentum, state.force, state.mass,
                              kwargs, scale_dev=s1)
    # chain.KE = KE(p) from the fused reduction, second half step (:660-665)
    s2 = _half_step(buf, buf, dof, kT_t, R, stepper.red(R))
    _lib.call('jmd_scale_momentum', _lib.dtype_code(R.dtype), P2.numel(),
              _lib.ptr(P2), _lib.ptr(s2), _lib.stream())
    return state.set(position=R2, momentum=P2, force=F2,
                     chain=_make_chain(buf, chain_length, chain.tau, dof))

  apply_fn._stepper = stepper
  return init_fn, apply_fn


# -- NPT Nose-Hoover --------------------------------------------------------

@dataclasses.dataclass
class NPTNoseHooverState:
  """simulate.py:704-760.  The box degrees of freedom are 0-d device tensors."""
  position: Any
  momentum: Any
  force: Any
  mass: Any
  reference_box: Any
  box_position: Any
  box_momentum: Any
  box_mass: Any
  dUdV: Any
  barostat: NoseHooverChain
  thermostat: NoseHooverChain

  @property
  def velocity(self):
    return self.momentum / self.mass

  @property
  def box(self):
    return npt_box(self)


def sinhx_x(x):
  """simulate.py:763-772: Taylor series of sinh(x) / x."""
  return (1 + x ** 2 / 6 + x ** 4 / 120 + x ** 6 / 5040 + x ** 8 / 362_880 + x ** 10 / 39_916_800)


def _box_volume(dim, box):
  # quantity.volume (quantity.py:111-121); the kernels serve diagonal matrices, whose
  # determinant is the product of the diagonal
  if box.ndim == 0:
    return box ** dim
  if box.ndim == 1:
    return torch.prod(box)
  if bool((torch.triu(box) == box).all()) or bool((torch.tril(box) == box).all()):
    return torch.prod(torch.diagonal(box))            # triangular: det without an LU
  return torch.linalg.det(box)


def _npt_box_info(state):
  """simulate.py:775-783."""
  dim = state.position.shape[1]
  ref = state.reference_box
  V_0 = _box_volume(dim, ref)
  V = V_0 * torch.exp(dim * state.box_position)
  return V, lambda V: (V / V_0) ** (1 / dim) * ref


def npt_box(state):
  """simulate.py:786-792."""
  V, box_fn = _npt_box_info(state)
  return box_fn(V)


def default_nhc_kwargs(tau, overrides):
  """simulate.py:520-536."""
  kw = dict(chain_length=3, chain_steps=2, sy_steps=3, tau=tau)
  kw.update(overrides or {})
  return kw


def _force_stress(energy_fn, R, box, kwargs):
  """simulate.py:848-855: (F, dU/d eps) of energy_fn(R, box=box, perturbation=1 + eps) at
  eps = 0.  The fused energies return both from ONE launch (the virial accumulated next to
  the forces); any other torch-differentiable energy goes through autograd like the
  reference."""
  nbr = kwargs.get('neighbor')
  fused = (getattr(energy_fn, '_jmd_fused', None) and hasattr(energy_fn, 'force_and_virial')
           and getattr(nbr, '_ws', None) is not None
           and not (hasattr(energy_fn, '_use_generic') and energy_fn._use_generic(nbr)))
  if fused:
    return energy_fn.force_and_virial(R, box=box, **kwargs)
  Rg = R.detach().requires_grad_(True)
  eps = torch.zeros((), dtype=R.dtype, device=R.device, requires_grad=True)
  E = energy_fn(Rg, box=box, perturbation=(1 + eps), **kwargs)
  gR, gE = torch.autograd.grad(E, (Rg, eps))
  return -gR, gE


def npt_nose_hoover(energy_fn, shift_fn, dt, pressure, kT, barostat_kwargs=None, thermostat_kwargs=None):
  """simulate.py:795-1004: NPT with a Nose-Hoover chain on the particles and one on the box
  (Tuckerman's direct translation).  Host-orchestrated: both chains run on the chain kernel
  (`jmd_nhc_half_step`; the barostat chain thermostats the single box momentum), forces AND
  dU/dV come from one launch of the fused force kernel (its virial accumulators), the
  exp(iL1) / exp(iL2) propagators are elementwise tensor arithmetic.  The new box reaches the
  kernels as a host value, which costs ONE device->host read per step (SURVEY 8f row 2:
  NPT is outside the graph-captured hot loop)."""
  dt_f = f32(dt)
  dt_2 = float(f32(dt / 2))
  dt = float(dt)
  bk = default_nhc_kwargs(1000 * dt, barostat_kwargs)
  tk = default_nhc_kwargs(100 * dt, thermostat_kwargs)
  for kw in (bk, tk):
    if kw['sy_steps'] not in SUZUKI_YOSHIDA_WEIGHTS:
      raise ValueError('sy_steps must be 1, 3, 5 or 7')

  def _kT_dev(_kT, R):
    if isinstance(_kT, torch.Tensor):
      return _kT.to(device=R.device, dtype=R.dtype).reshape(1)
    return torch.full((1,), float(_kT), dtype=R.dtype, device=R.device)

  def _half_step(kw, chain, buf_out, kT_t, R):
    scale = torch.empty((1,), dtype=R.dtype, device=R.device)
    _lib.call('jmd_nhc_half_step', _lib.dtype_code(R.dtype), kw['chain_length'], kw['chain_steps'],
              kw['sy_steps'], float(dt_f), float(f32(kw['tau'])), int(chain.degrees_of_freedom),
              _lib.ptr(kT_t), _lib.ptr(chain._buf), _lib.ptr(buf_out), None, _lib.ptr(scale), _lib.stream())
    return scale[0], _make_chain(buf_out, kw['chain_length'], chain.tau, chain.degrees_of_freedom)

  def _new_chain(kw, dof, KE, _kT, R):
    cl = kw['chain_length']
    buf = torch.zeros(3 * cl + 1, dtype=R.dtype, device=R.device)
    Q = f32(_kT) * (f32(kw['tau']) ** f32(2))                   # simulate.py:440-442
    buf[2 * cl:3 * cl] = float(Q)
    buf[2 * cl] = float(f32(Q * f32(dof)))
    buf[3 * cl] = KE
    return _make_chain(buf, cl, kw['tau'], dof)

  def _with_ke(chain, KE):
    buf = chain._buf.clone()
    buf[3 * chain._cl] = KE
    return _make_chain(buf, chain._cl, chain.tau, chain.degrees_of_freedom)

  def _box_mass(N, dim, _kT, tau, R):
    return torch.as_tensor(dim * (N + 1) * _kT * tau ** 2, dtype=R.dtype, device=R.device)

  def init_fn(key, R, box, mass=f32(1.0), momenta=None, **kwargs):
    _kT = kwargs.pop('kT', kT)
    _kT = float(_kT.detach().cpu()) if isinstance(_kT, torch.Tensor) else _kT
    R = R.contiguous()
    N, dim = R.shape
    zero = torch.zeros((), dtype=R.dtype, device=R.device)
    box_t = box if isinstance(box, torch.Tensor) else torch.as_tensor(np.asarray(box))
    box_t = box_t.to(device=R.device, dtype=R.dtype)
    if box_t.ndim == 0:
      box_t = torch.eye(dim, dtype=R.dtype, device=R.device) * box_t        # simulate.py:870-872
    force, dUdV = _force_stress(energy_fn, R, box_t, kwargs)
    m = _canonical_mass(mass, R)
    if momenta is None:
      P = initialize_momenta(R, m, key, _kT)
    else:
      P = momenta.to(dtype=R.dtype, device=R.device)
    P = P.contiguous()
    box_mass = _box_mass(N, dim, _kT, bk['tau'], R)
    KE = quantity.kinetic_energy(momentum=P, mass=m)
    return NPTNoseHooverState(R, P, force.contiguous(), m, box_t, zero, zero.clone(), box_mass, dUdV,
                              _new_chain(bk, 1, zero, _kT, R),
                              _new_chain(tk, quantity.count_dof(R), KE, _kT, R))

  def box_force(alpha, vol, dUdV, R, P, M, _pressure):
    dim = R.shape[1]
    KE2 = (P ** 2 / M).to(torch.float64).sum().to(R.dtype)      # util.high_precision_sum
    return alpha * KE2 - dUdV - _pressure * vol * dim

  def exp_iL1(box, R, V, V_b, **kwargs):
    x = V_b * dt
    x_2 = x / 2
    space_kw = {k: v for k, v in kwargs.items() if k != 'neighbor'}
    return shift_fn(R, R * (torch.exp(x) - 1) + dt * V * torch.exp(x_2) * sinhx_x(x_2), box=box, **space_kw)

  def exp_iL2(alpha, P, F, V_b):
    x = alpha * V_b * dt_2
    x_2 = x / 2
    return P * torch.exp(-x) + dt_2 * F * sinhx_x(x_2) * torch.exp(-x_2)

  def inner_step(state, **kwargs):
    _pressure = kwargs.pop('pressure', pressure)
    R, P, M, F = state.position, state.momentum, state.mass, state.force
    R_b, P_b, M_b = state.box_position, state.box_momentum, state.box_mass
    dUdV = state.dUdV                        # positions / box unchanged since the last evaluation
    N, dim = R.shape
    vol, box_fn = _npt_box_info(state)
    alpha = 1 + 1 / N
    G_e = box_force(alpha, vol, dUdV, R, P, M, _pressure)
    P_b = P_b + dt_2 * G_e
    P = exp_iL2(alpha, P, F, P_b / M_b)
    R_b = R_b + P_b / M_b * dt
    state = state.set(box_position=R_b)
    vol, box_fn = _npt_box_info(state)
    box = box_fn(vol)
    R = exp_iL1(box, R, P / M, P_b / M_b, **kwargs).contiguous()
    F, dUdV = _force_stress(energy_fn, R, box, kwargs)
    P = exp_iL2(alpha, P, F, P_b / M_b)
    G_e = box_force(alpha, vol, dUdV, R, P, M, _pressure)
    P_b = P_b + dt_2 * G_e
    return state.set(position=R, momentum=P.contiguous(), force=F, dUdV=dUdV, box_position=R_b,
                     box_momentum=P_b)

  def apply_fn(state, **kwargs):
    S = state
    _kT = kwargs.pop('kT', kT)
    R = S.position
    N, dim = R.shape
    kT_t = _kT_dev(_kT, R)
    kT_h = float(kT_t.cpu()) if isinstance(_kT, torch.Tensor) else _kT
    # update_mass of both chains happens inside the chain kernel (simulate.py:981-983)
    S = S.set(box_mass=_box_mass(N, dim, kT_h, S.barostat.tau, R))
    s_b, bc = _half_step(bk, S.barostat, torch.empty_like(S.barostat._buf), kT_t, R)
    s_t, tc = _half_step(tk, S.thermostat, torch.empty_like(S.thermostat._buf), kT_t, R)
    S = S.set(momentum=S.momentum * s_t, box_momentum=S.box_momentum * s_b)
    S = inner_step(S, **kwargs)
    tc = _with_ke(tc, quantity.kinetic_energy(momentum=S.momentum, mass=S.mass))
    bc = _with_ke(bc, quantity.kinetic_energy(momentum=S.box_momentum, mass=S.box_mass))
    s_t, tc = _half_step(tk, tc, tc._buf, kT_t, R)
    s_b, bc = _half_step(bk, bc, bc._buf, kT_t, R)
    return S.set(thermostat=tc, barostat=bc, momentum=(S.momentum * s_t).contiguous(),
                 box_momentum=S.box_momentum * s_b)

  return init_fn, apply_fn


def npt_nose_hoover_invariant(energy_fn, state, pressure, kT, **kwargs):
  """simulate.py:1007-1046."""
  volume, box_fn = _npt_box_info(state)
  PE = energy_fn(state.position, box=box_fn(volume), **kwargs)
  KE = kinetic_energy(state)
  DOF = quantity.count_dof(state.position)
  E = PE + KE
  c = state.thermostat
  E = E + c.momentum[0] ** 2 / (2 * c.mass[0]) + DOF * kT * c.position[0]
  for r, p, m in zip(c.position[1:], c.momentum[1:], c.mass[1:]):
    E = E + p ** 2 / (2 * m) + kT * r
  c = state.barostat
  for r, p, m in zip(c.position, c.momentum, c.mass):
    E = E + p ** 2 / (2 * m) + kT * r
  E = E + pressure * volume
  E = E + state.box_momentum ** 2 / (2 * state.box_mass)
  return E


# -- Langevin / Brownian ---------------------------------------------------

@dataclasses.dataclass
class NVTLangevinState:
  """simulate.py:1078-1099."""
  position: Any
  momentum: Any
  force: Any
  mass: Any
  rng: Any

  @property
  def velocity(self):
    return self.momentum / self.mass


def nvt_langevin(energy_or_force_fn, shift_fn, dt, kT, gamma=0.1, center_velocity=True, **sim_kwargs):
  """simulate.py:1113-1190: BAOAB Langevin.  B (kick) and A (drift) use the integrator
  kernels -- `jmd_nve_kick_drift` is called with half the drift length, twice -- the O step
  is elementwise tensor arithmetic around the device RNG (`torch.Generator`, the stand-in for
  the reference's PRNG key: streams differ, distributions agree)."""
  stepper = _Stepper(energy_or_force_fn, shift_fn, dt)

  def init_fn(key, R, mass=f32(1.0), momenta=None, **kwargs):
    _kT = kwargs.pop('kT', kT)
    R = R.contiguous()
    force = stepper.force(R, kwargs).contiguous()
    m = _canonical_mass(mass, R)
    g = _generator(key, R.device)
    P = initialize_momenta(R, m, g, _kT) if momenta is None else momenta.to(dtype=R.dtype, device=R.device)
    return NVTLangevinState(R, P.contiguous(), force, m, g)

  def step_fn(state, **kwargs):
    _dt = float(kwargs.pop('dt', dt))
    _kT = kwargs.pop('kT', kT)
    R, P, F, m, g = state.position, state.momentum, state.force, state.mass, state.rng
    dt_2 = _dt / 2
    space_kw = {k: v for k, v in kwargs.items() if k != 'neighbor'}
    P = P + dt_2 * F                                            # B
    R = stepper.shift_fn(R, dt_2 * P / m, **space_kw)           # A
    c1 = float(np.exp(-gamma * _dt))                            # O (simulate.py:1102-1110)
    c2 = torch.sqrt(torch.as_tensor(_kT, dtype=R.dtype, device=R.device) * (1 - c1 ** 2))
    P = c1 * P + c2 * torch.sqrt(m) * torch.randn(P.shape, dtype=P.dtype, device=P.device, generator=g)
    R = stepper.shift_fn(R, dt_2 * P / m, **space_kw).contiguous()    # A
    F = stepper.force(R, kwargs).contiguous()
    P = (P + dt_2 * F).contiguous()                             # B
    return NVTLangevinState(R, P, F, m, g)

  return init_fn, step_fn


@dataclasses.dataclass
class BrownianState:
  """simulate.py:1193-1208."""
  position: Any
  mass: Any
  rng: Any


def brownian(energy_or_force, shift, dt, kT, gamma=0.1):
  """simulate.py:1211-1268: overdamped Langevin dynamics."""
  force_fn = quantity.canonicalize_force(energy_or_force)

  def init_fn(key, R, mass=f32(1)):
    return BrownianState(R.contiguous(), _canonical_mass(mass, R), _generator(key, R.device))

  def apply_fn(state, **kwargs):
    _dt = kwargs.pop('dt', dt)
    _kT = kwargs.pop('kT', kT)
    _gamma = kwargs.pop('gamma', gamma)
    R, mass, g = state.position, state.mass, state.rng
    F = force_fn(R, **kwargs)
    xi = torch.randn(R.shape, dtype=R.dtype, device=R.device, generator=g)
    nu = 1.0 / (mass * _gamma)
    dR = F * _dt * nu + torch.sqrt(2.0 * _kT * _dt * nu) * xi
    space_kw = {k: v for k, v in kwargs.items() if k != 'neighbor'}
    return BrownianState(shift(R, dR, **space_kw).contiguous(), mass, g)

  return init_fn, apply_fn


def nvt_nose_hoover_invariant(energy_fn, state, kT, **kwargs):
  """simulate.py:672-701."""
  PE = energy_fn(state.position, **kwargs)
  KE = quantity.kinetic_energy(momentum=state.momentum, mass=state.mass)
  DOF = quantity.count_dof(state.position)
  E = PE + KE
  c = state.chain
  E = E + c.momentum[0] ** 2 / (2 * c.mass[0]) + DOF * kT * c.position[0]
  for r, p, m in zip(c.position[1:], c.momentum[1:], c.mass[1:]):
    E = E + p ** 2 / (2 * m) + kT * r
  return E


def kinetic_energy(state):
  return quantity.kinetic_energy(momentum=state.momentum, mass=state.mass)


def temperature(state):
  return quantity.temperature(momentum=state.momentum, mass=state.mass)
