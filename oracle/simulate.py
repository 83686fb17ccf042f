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
