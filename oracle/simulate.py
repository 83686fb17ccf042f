"""Oracle: NVE / NVT Nose-Hoover / FIRE steps (reference `jax_md/simulate.py`,
`jax_md/minimize.py`).  Test infrastructure only.  `force_fn(R) -> F`."""
import numpy as np

f32 = np.float32

SUZUKI_YOSHIDA_WEIGHTS = {   # simulate.py:328-347
    1: [1],
    3: [0.828981543588751, -0.657963087177502, 0.828981543588751],
    5: [0.2967324292201065, 0.2967324292201065, -0.186929716880426,
        0.2967324292201065, 0.2967324292201065],
    7: [0.784513610477560, 0.235573213359357, -1.17767998417887,
        1.31518632068391, -1.17767998417887, 0.235573213359357,
        0.784513610477560],
}


def kinetic_energy(P, mass):
  """quantity.py:124-159: 0.5 * high_precision_sum(p^2/m)."""
  return P.dtype.type(0.5 * (P ** 2 / mass).astype(np.float64).sum())


def temperature(P, mass):
  """quantity.py:162-199."""
  return P.dtype.type((P ** 2 / mass).astype(np.float64).sum() / P.size)


class State:
  def __init__(self, **kw):
    self.__dict__.update(kw)

  def copy(self, **kw):
    d = dict(self.__dict__)
    d.update(kw)
    return State(**d)


def velocity_verlet(force_fn, shift_fn, dt, state):
  """simulate.py:227-243."""
  dt = f32(dt)
  dt_2 = f32(dt / 2)
  P = state.momentum + dt_2 * state.force                     # :168-173
  R = shift_fn(state.position, dt * P / state.mass)           # :176-188
  F = force_fn(R)
  P = P + dt_2 * F
  return state.copy(position=R, momentum=P, force=F)


def nve(force_fn, shift_fn, dt=1e-3):
  """simulate.py:279-317."""
  def init_fn(R, momenta, mass=f32(1.0)):
    return State(position=R, momentum=momenta, force=force_fn(R), mass=mass)

  def apply_fn(state, dt_override=None):
    return velocity_verlet(force_fn, shift_fn,
                           dt if dt_override is None else dt_override, state)
  return init_fn, apply_fn


# -- Nose-Hoover chain (simulate.py:384-517) -----------------------------------

def nhc_init(dof, KE, kT, chain_length, tau, dtype):
  xi = np.zeros(chain_length, dtype)
  p_xi = np.zeros(chain_length, dtype)
  Q = kT * tau ** f32(2) * np.ones(chain_length, dtype=f32)
  Q[0] *= dof
  return State(position=xi, momentum=p_xi, mass=Q, tau=tau,
               kinetic_energy=KE, dof=dof)


def nhc_substep(delta, P, chain, kT, chain_length):
  """simulate.py:444-490."""
  xi, p_xi, Q, KE, DOF = (chain.position.copy(), chain.momentum.copy(),
                          chain.mass, chain.kinetic_energy, chain.dof)
  delta_2 = delta / f32(2.0)
  delta_4 = delta_2 / f32(2.0)
  delta_8 = delta_4 / f32(2.0)
  M = chain_length - 1
  old = p_xi.copy()
  G = old[M - 1] ** f32(2) / Q[M - 1] - kT
  p_xi[M] = old[M] + delta_4 * G
  p_new = p_xi[M]
  for m in range(M - 1, 0, -1):
    G = old[m - 1] ** 2 / Q[m - 1] - kT
    scale = np.exp(-delta_8 * p_new / Q[m + 1])
    p_new = scale * (scale * old[m] + delta_4 * G)
    p_xi[m] = p_new
  G = f32(2.0) * KE - DOF * kT
  scale = np.exp(-delta_8 * p_xi[1] / Q[1])
  p_xi[0] = scale * (scale * p_xi[0] + delta_4 * G)
  scale = np.exp(-delta_2 * p_xi[0] / Q[0])
  KE = KE * scale ** f32(2)
  P = P * scale
  xi = xi + delta_2 * p_xi / Q
  G = f32(2) * KE - DOF * kT
  for m in range(M):
    scale = np.exp(-delta_8 * p_xi[m + 1] / Q[m + 1])
    p_xi[m] = scale * (scale * p_xi[m] + delta_4 * G)
    G = p_xi[m] ** 2 / Q[m] - kT
  p_xi[M] = p_xi[M] + delta_4 * G
  return P, chain.copy(position=xi, momentum=p_xi, kinetic_energy=KE)


def nhc_half_step(P, chain, kT, dt, chain_length, chain_steps, sy_steps):
  """simulate.py:492-507."""
  if chain_steps == 1 and sy_steps == 1:
    return nhc_substep(dt, P, chain, kT, chain_length)
  delta = dt / chain_steps
  ws = SUZUKI_YOSHIDA_WEIGHTS[sy_steps]
  x64 = P.dtype == np.float64     # jnp.array(weights) is f64 only when x64 is on
  for i in range(chain_steps * sy_steps):
    w = ws[i % sy_steps]
    d = f32(np.float64(delta) * w) if x64 else f32(delta * f32(w))
    P, chain = nhc_substep(d, P, chain, kT, chain_length)
  return P, chain


def nvt_nose_hoover(force_fn, shift_fn, dt, kT, chain_length=5, chain_steps=2,
                    sy_steps=3, tau=None):
  """simulate.py:565-669."""
  dt = f32(dt)
  if tau is None:
    tau = dt * 100
  tau = f32(tau)

  def init_fn(R, momenta, mass=f32(1.0)):
    dof = R.size
    KE = kinetic_energy(momenta, mass)
    chain = nhc_init(dof, KE, kT, chain_length, tau, R.dtype)
    return State(position=R, momentum=momenta, force=force_fn(R), mass=mass,
                 chain=chain)

  def apply_fn(state, kT_override=None):
    _kT = kT if kT_override is None else kT_override
    chain = state.chain
    Q = _kT * tau ** f32(2) * np.ones(chain_length, dtype=f32)  # update_mass
    Q[0] *= chain.dof
    chain = chain.copy(mass=Q)
    P, chain = nhc_half_step(state.momentum, chain, _kT, dt, chain_length,
                             chain_steps, sy_steps)
    state = velocity_verlet(force_fn, shift_fn, dt, state.copy(momentum=P))
    chain = chain.copy(kinetic_energy=kinetic_energy(state.momentum, state.mass))
    P, chain = nhc_half_step(state.momentum, chain, _kT, dt, chain_length,
                             chain_steps, sy_steps)
    return state.copy(momentum=P, chain=chain)
  return init_fn, apply_fn


def nvt_nose_hoover_invariant(PE, state, kT):
  """simulate.py:672-701."""
  c = state.chain
  E = PE + kinetic_energy(state.momentum, state.mass)
  E += c.momentum[0] ** 2 / (2 * c.mass[0]) + c.dof * kT * c.position[0]
  for r, p, m in zip(c.position[1:], c.momentum[1:], c.mass[1:]):
    E += p ** 2 / (2 * m) + kT * r
  return E


# -- NPT Nose-Hoover (simulate.py:704-1046) --------------------------------------

def sinhx_x(x):
  """simulate.py:763-772."""
  return (1 + x ** 2 / 6 + x ** 4 / 120 + x ** 6 / 5040 + x ** 8 / 362_880 + x ** 10 / 39_916_800)


def volume(dim, box):
  """quantity.py:111-121 (matrix boxes here are diagonal: det = product of the diagonal)."""
  box = np.asarray(box)
  if box.ndim == 0:
    return box ** dim
  return np.prod(box) if box.ndim == 1 else np.prod(np.diag(box))


def _npt_box_info(state):
  """simulate.py:775-783."""
  dim = state.position.shape[1]
  ref = state.reference_box
  V_0 = volume(dim, ref)
  V = V_0 * np.exp(dim * state.box_position)
  return V, lambda V: (V / V_0) ** (1 / dim) * ref


def npt_box(state):
  """simulate.py:786-792."""
  V, box_fn = _npt_box_info(state)
  return box_fn(V)


def _nhc_update_mass(chain, kT, chain_length):
  Q = kT * chain.tau ** f32(2) * np.ones(chain_length, dtype=f32)     # simulate.py:509-515
  Q[0] *= chain.dof
  return chain.copy(mass=Q)


def npt_nose_hoover(force_stress_fn, shift_fn, dt, pressure, kT, barostat_kwargs=None,
                    thermostat_kwargs=None):
  """simulate.py:795-1004.  `force_stress_fn(R, box) -> (F, dUdV)` stands for the reference's
  value_and_grad over (position, eps) of `energy_fn(position, box=box, perturbation=1 + eps)`."""
  dt_2 = f32(dt / 2)

  def _kw(tau, over):
    d = dict(chain_length=3, chain_steps=2, sy_steps=3, tau=tau)       # simulate.py:520-536
    d.update(over or {})
    return d
  bk, tk = _kw(1000 * dt, barostat_kwargs), _kw(100 * dt, thermostat_kwargs)

  def _half(P, chain, _kT, kw):
    return nhc_half_step(P, chain, _kT, f32(dt), kw['chain_length'], kw['chain_steps'], kw['sy_steps'])

  def init_fn(R, box, momenta, mass=f32(1.0)):
    N, dim = R.shape
    zero, one = np.zeros((), R.dtype)[()], np.ones((), R.dtype)[()]
    box_mass = dim * (N + 1) * kT * bk['tau'] ** 2 * one
    KE_box = f32(0.5) * zero ** 2 / box_mass
    if np.ndim(box) == 0:
      box = np.eye(dim, dtype=np.asarray(box).dtype if isinstance(box, np.generic) else R.dtype) * box
    F, dUdV = force_stress_fn(R, box)
    KE = kinetic_energy(momenta, mass)
    return State(position=R, momentum=momenta, force=F, mass=mass, reference_box=box,
                 box_position=zero, box_momentum=zero, box_mass=box_mass, dUdV=dUdV,
                 barostat=nhc_init(1, KE_box, kT, bk['chain_length'], bk['tau'], R.dtype),
                 thermostat=nhc_init(R.size, KE, kT, tk['chain_length'], tk['tau'], R.dtype))

  def box_force(alpha, vol, dUdV, R, P, M, _pressure):
    N, dim = R.shape
    KE2 = np.sum((P ** 2 / M).astype(np.float64)).astype(R.dtype)
    return alpha * KE2 - dUdV - _pressure * vol * dim

  def exp_iL1(box, R, V, V_b):
    x = V_b * dt
    x_2 = x / 2
    return shift_fn(R, R * (np.exp(x) - 1) + dt * V * np.exp(x_2) * sinhx_x(x_2), box=box)

  def exp_iL2(alpha, P, F, V_b):
    x = alpha * V_b * dt_2
    x_2 = x / 2
    return P * np.exp(-x) + dt_2 * F * sinhx_x(x_2) * np.exp(-x_2)

  def inner_step(state, _pressure):
    R, P, M, F = state.position, state.momentum, state.mass, state.force
    R_b, P_b, M_b = state.box_position, state.box_momentum, state.box_mass
    dUdV = state.dUdV
    N, dim = R.shape
    vol, box_fn = _npt_box_info(state)
    alpha = 1 + 1 / N
    G_e = box_force(alpha, vol, dUdV, R, P, M, _pressure)
    P_b = P_b + dt_2 * G_e
    P = exp_iL2(alpha, P, F, P_b / M_b)
    R_b = R_b + P_b / M_b * dt
    state = state.copy(box_position=R_b)
    vol, box_fn = _npt_box_info(state)
    box = box_fn(vol)
    R = exp_iL1(box, R, P / M, P_b / M_b)
    F, dUdV = force_stress_fn(R, box)
    P = exp_iL2(alpha, P, F, P_b / M_b)
    G_e = box_force(alpha, vol, dUdV, R, P, M, _pressure)
    P_b = P_b + dt_2 * G_e
    return state.copy(position=R, momentum=P, force=F, dUdV=dUdV, box_position=R_b, box_momentum=P_b)

  def apply_fn(state, kT_override=None, pressure_override=None):
    S = state
    _kT = kT if kT_override is None else kT_override
    _p = pressure if pressure_override is None else pressure_override
    N, dim = S.position.shape
    bc = _nhc_update_mass(S.barostat, _kT, bk['chain_length'])
    tc = _nhc_update_mass(S.thermostat, _kT, tk['chain_length'])
    S = S.copy(box_mass=np.array(dim * (N + 1) * _kT * S.barostat.tau ** 2, S.position.dtype)[()])
    P_b, bc = _half(S.box_momentum, bc, _kT, bk)
    P, tc = _half(S.momentum, tc, _kT, tk)
    S = inner_step(S.copy(momentum=P, box_momentum=P_b), _p)
    tc = tc.copy(kinetic_energy=kinetic_energy(S.momentum, S.mass))
    bc = bc.copy(kinetic_energy=f32(0.5) * S.box_momentum ** 2 / S.box_mass)
    P, tc = _half(S.momentum, tc, _kT, tk)
    P_b, bc = _half(S.box_momentum, bc, _kT, bk)
    return S.copy(thermostat=tc, barostat=bc, momentum=P, box_momentum=P_b)
  return init_fn, apply_fn


def npt_nose_hoover_invariant(PE, state, pressure, kT):
  """simulate.py:1007-1046 with PE = energy_fn(position, box=npt_box(state))."""
  volume_, _ = _npt_box_info(state)
  E = PE + kinetic_energy(state.momentum, state.mass)
  c = state.thermostat
  E += c.momentum[0] ** 2 / (2 * c.mass[0]) + state.position.size * kT * c.position[0]
  for r, p, m in zip(c.position[1:], c.momentum[1:], c.mass[1:]):
    E += p ** 2 / (2 * m) + kT * r
  c = state.barostat
  for r, p, m in zip(c.position, c.momentum, c.mass):
    E += p ** 2 / (2 * m) + kT * r
  E += pressure * volume_
  E += state.box_momentum ** 2 / (2 * state.box_mass)
  return E


# -- FIRE (minimize.py:124-226) -----------------------------------------------

def fire_descent(force_fn, shift_fn, dt_start=0.1, dt_max=0.4, n_min=5,
                 f_inc=1.1, f_dec=0.5, alpha_start=0.1, f_alpha=0.99):
  def init_fn(R, mass=1.0):
    return State(position=R, momentum=np.zeros_like(R), force=force_fn(R),
                 mass=mass, dt=dt_start, alpha=alpha_start, n_pos=0)

  def apply_fn(state):
    state = velocity_verlet(force_fn, shift_fn, state.dt, state)
    R, P, F = state.position, state.momentum, state.force
    dt, alpha, n_pos = state.dt, state.alpha, state.n_pos
    F_norm = np.sqrt(np.sum(F ** 2) + 1e-6)
    P_norm = np.sqrt(np.sum(P ** 2))
    F_dot_P = np.sum(F * P)
    P = P + alpha * (F * P_norm / F_norm - P)
    n_pos = n_pos + 1 if F_dot_P >= 0 else 0
    if F_dot_P > 0:
      if n_pos > n_min:
        dt = min(dt * f_inc, dt_max)
        alpha = alpha * f_alpha
    if F_dot_P < 0:
      dt = dt * f_dec
      alpha = alpha_start
    P = (F_dot_P >= 0) * P
    return state.copy(momentum=P.astype(R.dtype), dt=dt, alpha=alpha,
                      n_pos=n_pos)
  return init_fn, apply_fn
