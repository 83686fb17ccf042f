"""Oracle: pair / Stillinger-Weber energies over neighbour lists.

Test infrastructure only.  Follows `jax_md/energy.py` (functional forms) and
`jax_md/smap.py:922-979` (`pair_neighbor_list` reduction rules).  The reference
gets forces by autodiff; the oracle differentiates the same expressions in
closed form and the tests check those against finite
differences of the oracle energy in f64.
"""
import numpy as np

from . import space
from .partition import Dense, Sparse, OrderedSparse, is_sparse, neighbor_list_mask

f32 = np.float32


# ----------------------------------------------------------------------------
# functional forms: each returns (U, dU/dr, dU/dsigma, dU/depsilon)
# ----------------------------------------------------------------------------

def lennard_jones(dr, sigma=1, epsilon=1, **unused):
  """energy.py:246-272."""
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    idr = sigma / dr
    idr = idr * idr
    idr6 = idr * idr * idr
    idr12 = idr6 * idr6
    return np.nan_to_num(f32(4) * epsilon * (idr12 - idr6))


def lennard_jones_grads(dr, sigma=1, epsilon=1):
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    x2 = (sigma / dr) ** 2
    x6 = x2 ** 3
    x12 = x6 * x6
    dU_dr = -24 * epsilon * (2 * x12 - x6) / dr
    dU_ds = 4 * epsilon * (12 * x12 - 6 * x6) / sigma
    dU_de = 4 * (x12 - x6) * np.ones_like(dr)
    return dU_dr, dU_ds, dU_de


def soft_sphere(dr, sigma=1, epsilon=1, alpha=2, **unused):
  """energy.py:125-173."""
  dr = dr / sigma
  with np.errstate(invalid='ignore'):
    base = np.where(dr < 1.0, f32(1.0) - dr, 0)
    fn = epsilon / alpha * base ** alpha
  return np.where(dr < 1.0, fn, f32(0.0)).astype(dr.dtype)


def soft_sphere_grads(dr, sigma=1, epsilon=1, alpha=2):
  x = dr / sigma
  inside = x < 1.0
  base = np.where(inside, 1.0 - x, 0)
  with np.errstate(invalid='ignore', divide='ignore'):
    bm1 = np.where(inside, base ** (alpha - 1), 0)
  dU_dr = np.where(inside, -(epsilon / sigma) * bm1, 0)
  dU_ds = np.where(inside, epsilon * bm1 * dr / sigma ** 2, 0)
  dU_de = np.where(inside, base ** alpha / alpha, 0)
  return dU_dr, dU_ds, dU_de


def morse(dr, sigma=1.0, epsilon=5.0, alpha=5.0, **unused):
  """energy.py:346-371."""
  with np.errstate(over='ignore', invalid='ignore'):
    U = epsilon * (f32(1) - np.exp(-alpha * (dr - sigma))) ** f32(2) - epsilon
  return np.nan_to_num(np.array(U, dtype=dr.dtype))


def morse_grads(dr, sigma=1.0, epsilon=5.0, alpha=5.0):
  m = np.exp(-alpha * (dr - sigma))
  dU_dr = 2 * epsilon * alpha * m * (1 - m)
  dU_ds = -dU_dr
  dU_de = (1 - m) ** 2 - 1
  return dU_dr, dU_ds, dU_de


def smooth_switch(dr, r_onset, r_cutoff):
  """energy.py:534-580 `smooth_fn` -> (S, dS/dr)."""
  r_c = r_cutoff ** f32(2)
  r_o = r_onset ** f32(2)
  r = dr ** f32(2)
  inner = np.where(dr < r_cutoff,
                   (r_c - r) ** 2 * (r_c + 2 * r - 3 * r_o) / (r_c - r_o) ** 3,
                   0)
  S = np.where(dr < r_onset, 1, inner)
  dinner = np.where(dr < r_cutoff,
                    12 * dr * (r_c - r) * (r_o - r) / (r_c - r_o) ** 3, 0)
  dS = np.where(dr < r_onset, 0, dinner)
  return S.astype(dr.dtype), dS.astype(dr.dtype)


class PairPotential:
  """U(dr; params) with optional multiplicative cutoff; params broadcast."""

  def __init__(self, kind, r_onset=None, r_cutoff=None):
    self.kind = kind
    self.r_onset = r_onset
    self.r_cutoff = r_cutoff
    self.fn, self.gfn = {
        'lj': (lennard_jones, lennard_jones_grads),
        'soft_sphere': (soft_sphere, soft_sphere_grads),
        'morse': (morse, morse_grads)}[kind]

  def energy(self, dr, **p):
    U = self.fn(dr, **p)
    if self.r_cutoff is not None:
      S, _ = smooth_switch(dr, self.r_onset, self.r_cutoff)
      U = S * U
    return U

  def grads(self, dr, **p):
    """(dU/dr, dU/dsigma, dU/depsilon) of the (switched) potential."""
    U = self.fn(dr, **p)
    with np.errstate(all='ignore'):
      g = [np.nan_to_num(x) for x in self.gfn(dr, **p)]
    if self.r_cutoff is not None:
      S, dS = smooth_switch(dr, self.r_onset, self.r_cutoff)
      g = [dS * U + S * g[0], S * g[1], S * g[2]]
    return g


# ----------------------------------------------------------------------------
# parameter expansion (smap.py:697-846)
# ----------------------------------------------------------------------------

def _expand(p, ia, ib, species):
  """Per-entry parameter for entries (a=row atom, b=neighbour)."""
  p_arr = np.asarray(p)
  if p_arr.ndim == 0:
    return p
  if species is None or p_arr.ndim == 1:
    if p_arr.ndim == 1:
      return 0.5 * (p_arr[ia] + p_arr[ib])          # default combinator :836
    if p_arr.ndim == 2:
      return p_arr[ia, ib]
    raise ValueError('bad parameter rank')
  if p_arr.ndim == 2:
    return p_arr[species[ia], species[ib]]
  raise ValueError('species parameters must be scalar or 2-d')


def _entries(R, nbrs):
  """(row atom a, neighbour b, mask) following smap.py:930-940."""
  N = R.shape[0]
  if is_sparse(nbrs.format):
    a, b = nbrs.idx[0], nbrs.idx[1]          # d(R[idx0], R[idx1])
    mask = nbrs.idx[0] < N
    return np.minimum(a, N - 1), np.minimum(b, N - 1), mask, None
  idx = nbrs.idx
  mask = idx < N
  rows = np.broadcast_to(np.arange(N)[:, None], idx.shape)
  # map_neighbor: displacement(R_neigh, R_i) -> a = neighbour, b = row atom
  return np.minimum(idx, N - 1), rows, mask, rows


def pair_neighbor_list_energy(pot, displacement, R, nbrs, species=None,
                              per_particle=False, want_grads=False, **params):
  """smap.py:922-979.  Returns E (or per-atom E); with want_grads also
  (force[N,dim], dE/dparam dict) by the chain rule on the same entries."""
  N, dim = R.shape
  a, b, mask, rows = _entries(R, nbrs)
  dR = displacement(R[a], R[b])
  dr = space.distance(dR)
  # parameter lookups: Sparse p[idx[0], idx[1]] (smap.py:712,722,794); Dense
  # p[row, idx] (smap.py:714,725,797) -- the ROW atom comes first there, while the
  # displacement is d(R_neigh, R_row) (map_neighbor)
  pa, pb = (b, a) if rows is not None else (a, b)
  p = {k: _expand(v, pa, pb, species) for k, v in params.items()}
  out = pot.energy(dr, **p) * mask
  norm = 1.0 if nbrs.format is OrderedSparse else 2.0
  if per_particle:
    if nbrs.format is OrderedSparse:
      raise ValueError('per-particle energies need Dense or Sparse')
    if is_sparse(nbrs.format):
      E = np.bincount(nbrs.idx[0][mask], weights=out[mask].astype(np.float64),
                      minlength=N)[:N]
    else:
      E = out.astype(np.float64).sum(axis=1)
    E = (E / norm).astype(R.dtype)
  else:
    E = R.dtype.type(out.astype(np.float64).sum() / norm)   # high_precision_sum
  if not want_grads:
    return E
  dU_dr, dU_ds, dU_de = pot.grads(dr, **p)
  with np.errstate(invalid='ignore', divide='ignore'):
    coef = np.where(mask & (dr > 0), dU_dr / np.where(dr > 0, dr, 1), 0) / norm
  g = (coef[..., None] * dR).astype(np.float64)
  grad = np.zeros((N, dim), np.float64)
  af, bf, gf = a.reshape(-1), b.reshape(-1), g.reshape(-1, dim)
  for k in range(dim):
    grad[:, k] += np.bincount(af, weights=gf[:, k], minlength=N)[:N]
    grad[:, k] -= np.bincount(bf, weights=gf[:, k], minlength=N)[:N]
  dparams = {}
  for name, dU in (('sigma', dU_ds), ('epsilon', dU_de)):
    if name not in params:
      continue
    w = (np.where(mask, dU, 0) / norm).astype(np.float64).reshape(-1)
    pv = np.asarray(params[name])
    if pv.ndim == 0:
      dparams[name] = w.sum()
    elif pv.ndim == 1:
      dparams[name] = 0.5 * (np.bincount(af, weights=w, minlength=N)[:N] +
                             np.bincount(bf, weights=w, minlength=N)[:N])
    else:
      paf, pbf = pa.reshape(-1), pb.reshape(-1)
      sa, sb = (species[paf], species[pbf]) if species is not None else (paf, pbf)
      out_t = np.zeros(pv.shape, np.float64)
      np.add.at(out_t, (sa, sb), w)
      dparams[name] = out_t
  return E, (-grad).astype(R.dtype), dparams


def pair_virial(pot, displacement, R, nbrs, species=None, **params):
  """dU/d(eps_ab) at eps = 0 for the box strain `perturbation=(I + eps)` of
  space.py:299-300: sum over list entries of (dU/dr)/r * dR_a * dR_b / norm -- what
  quantity.pressure / quantity.stress (quantity.py:202-282) obtain by autodiff."""
  N, dim = R.shape
  a, b, mask, rows = _entries(R, nbrs)
  dR = displacement(R[a], R[b])
  dr = space.distance(dR)
  pa, pb = (b, a) if rows is not None else (a, b)
  p = {k: _expand(v, pa, pb, species) for k, v in params.items()}
  dU_dr = pot.grads(dr, **p)[0]
  norm = 1.0 if nbrs.format is OrderedSparse else 2.0
  with np.errstate(invalid='ignore', divide='ignore'):
    coef = np.where(mask & (dr > 0), dU_dr / np.where(dr > 0, dr, 1), 0) / norm
  dRf = dR.reshape(-1, dim).astype(np.float64)
  return np.einsum('e,ea,eb->ab', coef.reshape(-1).astype(np.float64), dRf, dRf)


def pressure(pot, displacement, R, box, nbrs, kinetic_energy=0.0, species=None, **params):
  """quantity.py:202-235 for a scalar / vector box."""
  dim = R.shape[1]
  vol = float(box) ** dim if np.ndim(box) == 0 else float(np.prod(box))
  W = pair_virial(pot, displacement, R, nbrs, species=species, **params)
  return (2.0 * kinetic_energy - np.trace(W)) / (dim * vol)


def stress(pot, displacement, R, box, nbrs, mass=1.0, velocity=None, species=None, **params):
  """quantity.py:238-282."""
  dim = R.shape[1]
  vol = float(box) ** dim if np.ndim(box) == 0 else float(np.prod(box))
  W = pair_virial(pot, displacement, R, nbrs, species=species, **params)
  VxV = 0.0
  if velocity is not None:
    V = np.asarray(velocity, np.float64)
    VxV = np.einsum('n,na,nb->ab', np.broadcast_to(np.asarray(mass, np.float64), (len(V),)), V, V)
  return (VxV - W) / vol


def pair_energy_bruteforce(pot, displacement, R, species=None, **params):
  """smap.pair (smap.py:548-691) O(N^2) total energy, for cross-checks."""
  N = R.shape[0]
  ia, ib = np.meshgrid(np.arange(N), np.arange(N), indexing='ij')
  dR = displacement(R[ia], R[ib])
  dr = space.distance(dR)
  p = {k: _expand(v, ia, ib, species) for k, v in params.items()}
  U = pot.energy(dr, **p) * (ia != ib)
  return R.dtype.type(U.astype(np.float64).sum() / 2.0)


# ----------------------------------------------------------------------------
# Stillinger-Weber (energy.py:842-1014), Dense lists only
# ----------------------------------------------------------------------------

SW_DEFAULTS = dict(sigma=2.0951, A=7.049556277, B=0.6022245584, lam=21.0,
                   gamma=1.2, epsilon=2.16826, three_body_strength=1.0,
                   cutoff=3.77118)


def stillinger_weber_energy(displacement, R, nbrs, want_force=False, **kw):
  """energy.py:994-1012 with closed-form forces (optional)."""
  if nbrs.format is not Dense:
    raise NotImplementedError('Stillinger-Weber needs Dense neighbour lists.')
  p = dict(SW_DEFAULTS)
  p.update(kw)
  sigma, A, B, lam = p['sigma'], p['A'], p['B'], p['lam']
  gamma, eps, tbs, cutoff = (p['gamma'], p['epsilon'],
                             p['three_body_strength'], p['cutoff'])
  N, dim = R.shape
  idx = nbrs.idx
  mask = neighbor_list_mask(nbrs)
  j = np.minimum(idx, N - 1)
  dR = displacement(R[j], R[:, None, :])            # R_j - R_i, [N,M,dim]
  dr = space.distance(dR)
  a = cutoff / sigma

  # two body: energy.py:883-893
  within = (dr > 0) & (dr < cutoff)
  r = np.where(within, dr, 0)
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    term1 = B * (dr / sigma) ** (-4) - 1.0
    term2 = np.exp(1 / (r / sigma - a))
    two = np.where(within, term1 * term2, 0.0) * mask
  first = two.astype(np.float64).sum() / 2.0 * A

  # three body: energy.py:842-875 over the full M x M square
  d12 = np.where(dr < cutoff, dr, 0)                # [N,M]
  e1 = gamma / (d12 / sigma - a)
  # vmap order (energy.py:878-880): out[i, k, j] = f(dR12=dR[i,j], dR13=dR[i,k])
  t1 = np.exp(e1[:, None, :] + e1[:, :, None])
  nrm = dr + 1e-7                                    # quantity.py:285-289
  dot = np.einsum('ijd,ikd->ikj', dR, dR)
  cosang = np.clip(dot / nrm[:, None, :] / nrm[:, :, None], -1.0, 1.0)
  t2 = (cosang + 1.0 / 3) ** 2
  diff = dR[:, None, :, :] - dR[:, :, None, :]      # dR12 - dR13 at [i,k,j]
  sep = np.sqrt((diff ** 2).sum(-1))
  ok = (d12[:, None, :] > 0) & (d12[:, :, None] > 0) & (sep > 1e-5)
  mask_ijk = mask[:, None, :] * mask[:, :, None]
  three = np.where(ok, t1 * t2, 0) * mask_ijk
  second = lam * three.astype(np.float64).sum() / 2.0
  E = R.dtype.type(eps * (first + tbs * second))
  if not want_force:
    return E

  # ---- forces: differentiate the expressions above --------------------------
  grad = np.zeros((N, dim), np.float64)
  rows = np.broadcast_to(np.arange(N)[:, None], idx.shape)
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    x = r / sigma
    df2 = np.where(
        within,
        (-4 * B * sigma ** 4 * dr ** -5.0) * term2
        - term1 * term2 / (sigma * (x - a) ** 2), 0.0) * mask
    unit = np.where(dr[..., None] > 0, dR / np.where(dr > 0, dr, 1)[..., None], 0)
  g2 = (eps * A / 2.0) * df2[..., None] * unit       # d/d(R_j) ; -that on R_i
  # three-body: V = sum_{j,k} h(r_ij) h(r_ik) (c + 1/3)^2 / 2 * eps*lam*tbs
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    inside = (d12 > 0)
    dh = np.where(inside, -gamma / (sigma * (d12 / sigma - a) ** 2), 0)  # dlog h/dr
  pref = eps * lam * tbs / 2.0
  w = np.where(ok, t1, 0) * mask_ijk                 # [i,k,j]
  cterm = cosang + 1.0 / 3
  clipped = (dot / nrm[:, None, :] / nrm[:, :, None])
  live = (np.abs(clipped) <= 1.0)
  # d/d r_ij (radial, through h and through the 1/(r+1e-7) in cos)
  # index names: axis1 = k (dR13), axis2 = j (dR12)
  rj = nrm[:, None, :]
  rk = nrm[:, :, None]
  uj = unit[:, None, :, :]
  uk = unit[:, :, None, :]
  dRj = dR[:, None, :, :]
  dRk = dR[:, :, None, :]
  # dcos/d(dR_j) = dR_k/(rj rk) - cos * u_j / rj     (u_j = dR_j/|dR_j|)
  dcos_dj = (dRk / (rj * rk)[..., None]
             - (clipped / rj)[..., None] * uj) * live[..., None]
  dcos_dk = (dRj / (rj * rk)[..., None]
             - (clipped / rk)[..., None] * uk) * live[..., None]
  common = pref * w
  gj = (common * t2 * dh[:, None, :])[..., None] * uj \
      + (common * 2 * cterm)[..., None] * dcos_dj     # d/d(dR_ij)
  gk = (common * t2 * dh[:, :, None])[..., None] * uk \
      + (common * 2 * cterm)[..., None] * dcos_dk     # d/d(dR_ik)
  Gj = gj.sum(axis=1) + g2                            # [i, j, dim]
  Gk = gk.sum(axis=2)                                 # [i, k, dim]
  Gtot = (Gj + Gk).astype(np.float64)                 # d/d(dR_i,slot)
  jf = j.reshape(-1)
  rf = rows.reshape(-1)
  Gf = Gtot.reshape(-1, dim)
  for k in range(dim):
    grad[:, k] += np.bincount(jf, weights=Gf[:, k], minlength=N)[:N]
    grad[:, k] -= np.bincount(rf, weights=Gf[:, k], minlength=N)[:N]
  return E, (-grad).astype(R.dtype)
