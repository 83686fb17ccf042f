"""`smap.pair_neighbor_list`: drop-in for jax_md/smap.py:856-979 for the pair
potentials of this package (Lennard-Jones, soft sphere, Morse, optionally
wrapped in `energy.multiplicative_isotropic_cutoff`).

The mapped function has the reference signature
`fn_mapped(R, neighbor, **dynamic_kwargs)` and returns the total energy (or
per-atom energies with `reduce_axis=(1,)`).  Instead of gather + vmap + sum it
launches the fused CUDA kernel over the neighbour list's internal full rows.
Gradients with respect to positions and to `sigma` / `epsilon` are served by a
custom backward (the reference's `jax.custom_vjp` role is played by
`torch.autograd.Function`): the forward kernel already produced
F = -dE/dR, dE/dsigma and dE/depsilon.  Higher-order derivatives are out of
scope.
"""
import ctypes as C
import enum
from typing import Any

import numpy as np
import torch

from . import _lib, partition, space
from . import dataclasses as _dataclasses

_DEFAULTS = {
    _lib.POT_LJ: {'sigma': 1.0, 'epsilon': 1.0, 'alpha': 1.0},
    _lib.POT_SOFT_SPHERE: {'sigma': 1.0, 'epsilon': 1.0, 'alpha': 2.0},
    _lib.POT_MORSE: {'sigma': 1.0, 'epsilon': 5.0, 'alpha': 5.0},
}
_PARAM_ORDER = ('sigma', 'epsilon', 'alpha')


class Scratch:
  """Per-(device, n) reduction scratch shared by the force kernels."""
  _cache = {}

  @classmethod
  def get(cls, n, device):
    key = (int(n), str(device))
    s = cls._cache.get(key)
    if s is None:
      nd = _lib.load().jmd_red_scratch_doubles(int(n))
      s = torch.zeros(nd, dtype=torch.float64, device=device)
      cls._cache[key] = s
    return s


class ParameterTreeMapping(enum.Enum):
  """smap.py:57-79."""
  Global = 0
  PerParticle = 1
  PerBond = 2
  PerSpecies = 3


@_dataclasses.dataclass
class ParameterTree:
  """smap.py:82-93: parameters in the form of a tree (dict / list / tuple of arrays),
  processed leaf by leaf according to `mapping`."""
  tree: Any
  mapping: ParameterTreeMapping = _dataclasses.static_field()


def _tree_map(f, tree):
  if isinstance(tree, dict):
    return {k: _tree_map(f, v) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(_tree_map(f, v) for v in tree)
  return f(tree)


def _needs_generic(v):
  """Parameter forms only the generic (torch-composed) path serves: ParameterTrees and
  `(combinator, per_atom_array)` pairs (smap.py:826-846)."""
  return isinstance(v, ParameterTree) or (isinstance(v, tuple) and len(v) == 2 and callable(v[0]))


def _merge(static, dynamic, ignore_unused):
  """util.merge_dicts (util.py:57-79)."""
  if not ignore_unused:
    return {**static, **dynamic}
  merged = dict(static)
  for k in static:
    if dynamic.get(k) is not None:
      merged[k] = dynamic[k]
  return merged


def pair_descriptor(pot, sigma=None, epsilon=None, alpha=None):
  """`_lib.PairT` for scalar parameters (host arithmetic only; shared with the XLA-FFI
  binding): potential kind, cutoff constants in the reference's dtype (energy.py:558-566)."""
  pt = _lib.PairT()
  kind = pot['kind']
  pt.kind = kind
  pt.has_cutoff = 1 if pot.get('r_cutoff') is not None else 0
  if pt.has_cutoff:
    pt.r_onset = float(pot['r_onset'])
    pt.r_cutoff = float(pot['r_cutoff'])
    r_o = pot['r_onset'] ** np.float32(2)
    r_c = pot['r_cutoff'] ** np.float32(2)
    pt.r_onset2 = float(r_o)
    pt.r_cutoff2 = float(r_c)
    pt.switch_denom = float((r_c - r_o) ** 3)
  for k, (name, v) in enumerate(zip(_PARAM_ORDER, (sigma, epsilon, alpha))):
    pt.mode[k] = _lib.PARAM_SCALAR
    pt.scalar[k] = float(_DEFAULTS[kind][name] if v is None else v)
  return pt


class PairNeighborListFn:
  """Callable returned by `pair_neighbor_list`."""
  _jmd_fused = 'pair'

  def __init__(self, pot, displacement_or_metric, species, reduce_axis,
               ignore_unused_parameters, kwargs, fn=None):
    # the same potential through the generic, torch-composed path: lists whose `idx` is
    # not the kernel's own (custom_mask_function, a foreign NeighborList), asymmetric
    # parameter tables, per-edge output (reduce_axis=())
    self.generic = None if fn is None else GenericPairNeighborListFn(
        fn, displacement_or_metric, species, reduce_axis, ignore_unused_parameters, kwargs)
    self.pot = pot                      # dict(kind, r_onset, r_cutoff)
    self.spec = space.get_spec(displacement_or_metric)
    self.species = species
    self.reduce_axis = reduce_axis
    self.ignore_unused = ignore_unused_parameters
    self.kwargs = dict(kwargs)
    self.kwargs.pop('fractional_coordinates', None)   # energy.py:240,340,443
    self._conv = {}
    self.always_generic = False
    if reduce_axis is not None:
      if len(reduce_axis) == 0:
        self.always_generic = True        # per-edge output (smap.py:952-954)
      elif 0 in reduce_axis and 1 not in reduce_axis:
        raise ValueError()
    for v in self.kwargs.values():        # asymmetric [S, S] / [N, N] tables: row forces are not -dE/dR_i
      if isinstance(v, (torch.Tensor, np.ndarray)) and getattr(v, 'ndim', 0) == 2:
        t = torch.as_tensor(v)
        if t.dtype.is_floating_point and not bool(torch.equal(t, t.T)):
          self.always_generic = True
    if self.always_generic and self.generic is None:
      raise NotImplementedError('this parameter form needs the generic pair_neighbor_list path')

  def _use_generic(self, neighbor):
    if self.generic is None:
      return False
    if self.always_generic:
      return True
    return getattr(neighbor, '_ws', None) is None or not neighbor.internal_list_is_current

  # -- parameter canonicalisation (smap.py:697-846) ---------------------------
  def _tensor(self, x, dtype, device):
    # converted copies are cached per tensor object AND version (in-place edits make a
    # new entry); NumPy inputs can be mutated without notice, so they are never cached
    # (small tables are keyed by their bytes instead: the usual [S, S] species tables then
    #  convert once, which also keeps the call capturable in a CUDA graph)
    cacheable = isinstance(x, torch.Tensor)
    if cacheable:
      key = (id(x), dtype, x._version)
    else:
      a = np.ascontiguousarray(x)
      key = ('np', a.shape, a.dtype.str, a.tobytes(), dtype, str(device)) if a.size <= 4096 else None
    hit = self._conv.get(key) if key is not None else None
    if hit is not None and (hit[0] is x or not cacheable):
      return hit[1]
    if isinstance(x, torch.Tensor):
      t = x.detach().to(device=device, dtype=dtype).contiguous()
    else:
      t = torch.as_tensor(np.asarray(x), dtype=dtype, device=device).contiguous()
    if t.ndim == 2 and t.dtype.is_floating_point and not bool(torch.equal(t, t.T)):
      # The fused kernel takes the force on atom i from row i alone, which is
      # -dE/dR_i only when p[a, b] == p[b, a]; checked once per parameter object.
      raise NotImplementedError(
          'asymmetric [S, S] / [N, N] parameter tables are not served by the fused '
          'kernel; wrap the potential in a plain Python function to take the generic '
          'smap.pair_neighbor_list path.')
    if len(self._conv) > 64:
      self._conv.clear()
    if key is not None:
      self._conv[key] = (x if cacheable else None, t)
    return t

  def _pair_struct(self, R, species, params, sparse=False):
    pt = _lib.PairT()
    pt.transposed = 1 if sparse else 0
    keep = []
    kind = self.pot['kind']
    pt.kind = kind
    pt.has_cutoff = 1 if self.pot.get('r_cutoff') is not None else 0
    if pt.has_cutoff:
      pt.r_onset = float(self.pot['r_onset'])
      pt.r_cutoff = float(self.pot['r_cutoff'])
      # energy.py:558-559: r_c = r_cutoff ** f32(2), r_o = r_onset ** f32(2),
      # evaluated in the dtype the factory left them in (f32 unless f64 arrays)
      r_o = self.pot['r_onset'] ** np.float32(2)
      r_c = self.pot['r_cutoff'] ** np.float32(2)
      pt.r_onset2 = float(r_o)
      pt.r_cutoff2 = float(r_c)
      pt.switch_denom = float((r_c - r_o) ** 3)
    pt.n_species = 0
    modes = []
    for k, name in enumerate(_PARAM_ORDER):
      v = params.get(name, _DEFAULTS[kind][name])
      ndim = v.ndim if isinstance(v, (torch.Tensor, np.ndarray)) else 0
      if ndim == 0:
        pt.mode[k] = _lib.PARAM_SCALAR
        pt.scalar[k] = float(v)
      elif species is None or ndim == 1:
        if ndim == 1:
          pt.mode[k] = _lib.PARAM_PER_ATOM
        elif ndim == 2:
          pt.mode[k] = _lib.PARAM_MATRIX
          pt.n_species = int(v.shape[0])
        else:
          raise ValueError('Parameter array must be either a scalar, a vector, '
                           f'or a matrix. Found ndim={ndim}.')
        t = self._tensor(v, R.dtype, R.device)
        keep.append(t)
        pt.array[k] = t.data_ptr()
      else:
        if ndim != 2:
          raise ValueError('Params must be a scalar or a 2d array if using a '
                           'species lookup.')
        pt.mode[k] = _lib.PARAM_SPECIES
        pt.n_species = int(v.shape[0])
        t = self._tensor(v, R.dtype, R.device)
        keep.append(t)
        pt.array[k] = t.data_ptr()
      modes.append(pt.mode[k])
    return pt, keep, modes

  def _resolve(self, neighbor, dynamic_kwargs):
    if neighbor is None:
      neighbor = dynamic_kwargs.pop('neighbor', None)
    if neighbor is None:
      raise TypeError('energy_fn(R, neighbor=...) needs a NeighborList')
    if neighbor._ws is None or not neighbor.internal_list_is_current:
      raise NotImplementedError(
          'The fused kernels read the internal rows of a NeighborList built by '
          'jax_md_b200.partition; this list has a foreign `idx`.')
    species = dynamic_kwargs.pop('species', self.species)
    perturbation = dynamic_kwargs.pop('perturbation', self.kwargs.get('perturbation'))
    neighbor._ws.set_box(self.spec, dynamic_kwargs.pop('box', None))     # periodic_general: box override
    params = _merge(self.kwargs, dynamic_kwargs, self.ignore_unused)
    params.pop('perturbation', None)
    if perturbation is not None:
      # space.py:299-300 scales the displacement; the fused kernel has no such input.
      # Its derivative at the identity (all the reference uses it for) is `virial()`.
      raise NotImplementedError(
          'perturbation= is not an input of the fused kernel: use quantity.pressure / '
          'quantity.stress (served from its virial) or a plain Python potential (generic '
          'smap.pair_neighbor_list path).')
    return neighbor, species, params

  def _species_tensor(self, species, device):
    if species is None:
      return None
    if isinstance(species, torch.Tensor) and species.dtype == torch.int32 \
        and species.is_cuda and species.is_contiguous():
      return species
    return self._tensor(species, torch.int32, device)

  # -- kernels ------------------------------------------------------------------
  def launch(self, R, neighbor, species, params, want_energy, per_atom=False,
             momentum=None, mass=None, dt_2=0.0, dt_dev=None, red=None,
             refresh_positions=True):
    """Runs the fused kernel.  Returns dict(force, red, e_atom, dparam, modes)."""
    ws = neighbor._ws
    dev, N, dim = R.device, ws.n, ws.dim
    ws.set_species(self._species_tensor(species, dev))
    if refresh_positions:
      _lib.call('jmd_nbr_pack', ws.ref(), _lib.ptr(R.contiguous()), _lib.stream())
    pt, keep, modes = self._pair_struct(
        R, species, params, partition.is_sparse(neighbor.format))
    force = torch.empty((N, dim), dtype=R.dtype, device=dev)
    if red is None:
      red = torch.zeros(_lib.RED_COUNT, dtype=torch.float64, device=dev)
    partials = Scratch.get(N, dev)
    e_atom = torch.empty((N,), dtype=R.dtype, device=dev) if per_atom else None
    dparam = None
    pt.dparam_rows = 0
    if want_energy and any(m in (_lib.PARAM_SPECIES, _lib.PARAM_PER_ATOM)
                           for m in modes[:2]):
      size = 2 * N if _lib.PARAM_PER_ATOM in modes[:2] else 2 * pt.n_species ** 2
      size = max(size, 2 * pt.n_species ** 2)
      if _lib.PARAM_PER_ATOM not in modes[:2] and 1 <= pt.n_species <= _lib.DPARAM_MAX_SPECIES:
        # per-atom rows instead of atomics into the [S, S] tables: reproducible sums
        pt.dparam_rows = 1
        size = 2 * N * pt.n_species
      dparam = torch.zeros(size, dtype=torch.float64, device=dev)
    mass_is_array = 0
    if momentum is not None:
      mass_is_array = 1 if mass.numel() > 1 else 0
    _lib.call('jmd_pair_force', ws.ref(), C.byref(pt), _lib.ptr(force),
              _lib.ptr(e_atom), _lib.ptr(red), _lib.ptr(dparam),
              _lib.ptr(partials), _lib.ptr(momentum), _lib.ptr(mass),
              mass_is_array, float(dt_2), _lib.ptr(dt_dev),
              1 if want_energy else 0, _lib.stream())
    if pt.dparam_rows:
      dparam = self._fold_species_rows(dparam, ws, N, pt.n_species, bool(pt.transposed))
    return dict(force=force, red=red, e_atom=e_atom, dparam=dparam, modes=modes,
                n_species=pt.n_species, keep=keep)

  @staticmethod
  def _fold_species_rows(rows, ws, N, S, transposed):
    """[2, N, S] per-atom sums (home atom i, neighbour species) -> the two [S, S] tables the
    autograd function reads: table[s_i] = sum of the rows of species s_i.  One f64 GEMM with the
    one-hot species matrix: a fixed summation order, unlike atomics."""
    species = ws.species
    onehot = torch.zeros((N, S), dtype=torch.float64, device=rows.device)
    onehot.scatter_(1, species[:N].long().reshape(-1, 1), 1.0)
    tables = torch.matmul(onehot.T.unsqueeze(0), rows.view(2, N, S))      # [2, S, S]
    if transposed:
      tables = tables.transpose(1, 2)
    return tables.contiguous().reshape(-1)

  def force_and_virial(self, R, neighbor=None, **dynamic_kwargs):
    """One launch -> (force [N, dim], trace of `virial()` as a 0-d tensor): what
    simulate.npt_nose_hoover's force_stress_fn (simulate.py:848-855) differentiates for."""
    neighbor, species, params = self._resolve(neighbor, dict(dynamic_kwargs))
    out = self.launch(R, neighbor, species, params, True)
    v = out['red'][_lib.RED_VIRIAL:_lib.RED_VIRIAL + R.shape[1]]
    return out['force'], v.sum().to(R.dtype)

  def virial(self, R, neighbor=None, **dynamic_kwargs):
    """dU/d(eps_ab) at eps = 0 for the box perturbation `(I + eps)` of
    space.py:299-300, i.e. sum over pairs of (dU/dr)/r * d_a * d_b -- the
    quantity `quantity.pressure` / `quantity.stress` obtain in the reference by
    differentiating through `perturbation=` (quantity.py:226-235, 268-282);
    here it is a by-product of the fused force kernel.  -> [dim, dim]."""
    neighbor, species, params = self._resolve(neighbor, dict(dynamic_kwargs))
    red = self.launch(R, neighbor, species, params, True)['red']
    v = red[_lib.RED_VIRIAL:_lib.RED_VIRIAL + 6].to(R.dtype)      # xx yy zz xy xz yz
    dim = R.shape[1]
    if dim == 2:
      return torch.stack([torch.stack([v[0], v[3]]), torch.stack([v[3], v[1]])])
    return torch.stack([torch.stack([v[0], v[3], v[4]]),
                        torch.stack([v[3], v[1], v[5]]),
                        torch.stack([v[4], v[5], v[2]])])

  def force(self, R, neighbor=None, **dynamic_kwargs):
    """-dE/dR straight from the kernel (quantity.force fast path)."""
    if neighbor is None:
      neighbor = dynamic_kwargs.pop('neighbor', None)
    if self._use_generic(neighbor):
      Rg = R.detach().requires_grad_(True)
      with torch.enable_grad():
        (g,) = torch.autograd.grad(self.generic(Rg, neighbor, **dynamic_kwargs), Rg)
      return -g
    neighbor, species, params = self._resolve(neighbor, dict(dynamic_kwargs))
    return self.launch(R, neighbor, species, params, want_energy=False)['force']

  def __call__(self, R, neighbor=None, **dynamic_kwargs):
    if neighbor is None:
      neighbor = dynamic_kwargs.pop('neighbor', None)
    if self._use_generic(neighbor):
      return self.generic(R, neighbor, **dynamic_kwargs)
    neighbor, species, params = self._resolve(neighbor, dict(dynamic_kwargs))
    per_atom = self.reduce_axis is not None
    if per_atom and neighbor.format is partition.OrderedSparse:
      raise ValueError('Cannot report per-particle values with a neighbor list '
                       'whose format is `OrderedSparse`. Please use either '
                       '`Dense` or `Sparse`.')
    grad_params = [(n, params[n]) for n in ('sigma', 'epsilon')
                   if isinstance(params.get(n), torch.Tensor)
                   and params[n].requires_grad]
    if torch.is_grad_enabled():
      for n, v in params.items():
        if n not in ('sigma', 'epsilon') and isinstance(v, torch.Tensor) and v.requires_grad:
          raise NotImplementedError(
              f'd(energy)/d({n}) is not produced by the fused kernel (only positions, sigma and '
              'epsilon are); use a plain Python potential for the generic smap.pair_neighbor_list path.')
    needs_grad = torch.is_grad_enabled() and (R.requires_grad or grad_params)
    if not needs_grad:
      out = self.launch(R, neighbor, species, params, True, per_atom)
      if per_atom:
        return out['e_atom']
      return out['red'][_lib.RED_ENERGY].to(R.dtype)
    if per_atom:
      raise NotImplementedError('gradients of per-particle energies')
    fn = self

    class _Energy(torch.autograd.Function):
      @staticmethod
      def forward(ctx, Rin, *ptensors):
        out = fn.launch(Rin.detach(), neighbor, species, params, True, False)
        ctx.out = out
        ctx.shapes = [p.shape for p in ptensors]
        return out['red'][_lib.RED_ENERGY].to(Rin.dtype)

      @staticmethod
      def backward(ctx, g):
        out = ctx.out
        grads = [-(g * out['force'])]
        for (name, p), k in zip(grad_params, range(len(grad_params))):
          slot = 0 if name == 'sigma' else 1
          mode = out['modes'][slot]
          if mode == _lib.PARAM_SCALAR:
            r = out['red'][_lib.RED_DSIGMA + slot]
            grads.append((g * r).to(p.dtype).reshape(p.shape))
          elif mode == _lib.PARAM_SPECIES:
            S = out['n_species']
            tbl = out['dparam'][slot * S * S:(slot + 1) * S * S].reshape(S, S)
            grads.append((g * tbl).to(p.dtype))
          elif mode == _lib.PARAM_PER_ATOM:
            N = Rin_n[0]
            grads.append((g * out['dparam'][slot * N:(slot + 1) * N]).to(p.dtype))
          else:
            raise NotImplementedError('gradient of [N, N] matrix parameters')
        return tuple(grads)

    Rin_n = [R.shape[0]]
    return _Energy.apply(R, *[p for _, p in grad_params])


class GenericPairNeighborListFn:
  """`smap.pair_neighbor_list` (smap.py:856-979) for an ARBITRARY Python
  `fn(dr, **params)` written with torch ops -- SURVEY.md 8(f) row 1.  The
  neighbour list comes from the CUDA builder; the energy is composed exactly as
  the reference composes it (gather `R[idx]` with the padding index clamped,
  metric, `fn`, mask `idx < N`, high-precision sum / normalisation) and is
  differentiable through `torch.autograd` (forces via `quantity.force`).  This
  is the generic, un-fused path: the potentials of `energy.py` never take it."""
  _jmd_fused = None

  def __init__(self, fn, displacement_or_metric, species, reduce_axis,
               ignore_unused_parameters, kwargs):
    self.fn = fn
    self.d = displacement_or_metric
    self.species = species
    self.reduce_axis = reduce_axis
    self.ignore_unused = ignore_unused_parameters
    self.kwargs = dict(kwargs)
    self.kwargs.pop('fractional_coordinates', None)
    self._is_metric = None

  def _metric(self, Ra, Rb, **kw):
    """space.canonicalize_displacement_or_metric (space.py:505-522)."""
    if self._is_metric is None:
      probe = self.d(Ra.reshape(-1, Ra.shape[-1])[:1], Rb.reshape(-1, Rb.shape[-1])[:1])
      self._is_metric = probe.ndim == 1
    out = self.d(Ra, Rb, **kw)
    return out if self._is_metric else space.distance(out)

  def _param(self, v, i, j, si, sj, R):
    """smap.py:697-846: scalar | per-atom [N] through a combinator (default
    mean) | [N, N] | species table [S, S]; `(combinator, per_atom)` tuples."""
    comb = lambda a, b: 0.5 * (a + b)
    if isinstance(v, tuple) and len(v) == 2 and callable(v[0]):
      comb, v = v
    if isinstance(v, ParameterTree):                      # smap.py:732-768, 806-823
      M = ParameterTreeMapping
      leaf = lambda p: p if isinstance(p, torch.Tensor) else torch.as_tensor(p, device=R.device)
      if v.mapping is M.Global:
        return v.tree
      if si is not None:
        if v.mapping is not M.PerSpecies:
          raise ValueError('Parameter tree mapping must be either Global or PerSpecies if using '
                           'a species lookup.')
        return _tree_map(lambda p: leaf(p)[si, sj], v.tree)
      if v.mapping is M.PerParticle:
        return _tree_map(lambda p: comb(leaf(p)[i], leaf(p)[j]), v.tree)
      if v.mapping is M.PerBond:
        return _tree_map(lambda p: leaf(p)[i, j], v.tree)
      raise ValueError('Without species information ParameterTreeMapping be Global or PerParticle. '
                       f'Found {v.mapping}.')
    if isinstance(v, (int, float)):
      return v
    v = torch.as_tensor(v, device=R.device) if not isinstance(v, torch.Tensor) else v.to(R.device)
    if v.ndim == 0:
      return v
    if si is not None:
      if v.ndim == 2:
        return v[si, sj]
      raise ValueError('Parameters with species must be scalars or [S, S] tables; found '
                       f'shape {tuple(v.shape)}.')
    if v.ndim == 1:
      return comb(v[i], v[j])
    if v.ndim == 2:
      return v[i, j]
    raise ValueError(f'Parameter must be a scalar, a vector or a matrix; found {v.ndim} dims.')

  def __call__(self, R, neighbor=None, **dynamic_kwargs):
    if neighbor is None:
      raise ValueError('neighbor must be passed (positionally or as neighbor=).')
    N = R.shape[0]
    species = dynamic_kwargs.pop('species', self.species)
    merged = _merge(self.kwargs, dynamic_kwargs, self.ignore_unused)
    space_kw = {k: merged.pop(k) for k in ('perturbation',) if k in merged}
    idx = neighbor.idx.long()
    sparse = partition.is_sparse(neighbor.format)
    if sparse:
      recv = idx[0]
      mask = recv < N
      # entries are (idx[0], idx[1]) in that order for the metric AND the parameter
      # lookups p[idx[0], idx[1]] (smap.py:712-725, 935-937)
      i, j = idx[0].clamp(max=N - 1), idx[1].clamp(max=N - 1)
      dr = self._metric(R[i], R[j], **space_kw)
      norm = 1.0 if neighbor.format is partition.OrderedSparse else 2.0
    else:
      mask = idx < N
      j = idx.clamp(max=N - 1)                                   # OOB gather clamps
      i = torch.arange(N, device=R.device)[:, None].expand_as(j)
      dr = self._metric(R[:, None, :], R[j], **space_kw)        # smap.py:938-940
      norm = 2.0
    si = sj = None
    if species is not None:
      sp = torch.as_tensor(species, device=R.device).long()
      si, sj = sp[i], sp[j]
    params = {k: self._param(v, i, j, si, sj, R) for k, v in merged.items()}
    out = self.fn(dr, **params)
    out = torch.where(mask, out, torch.zeros_like(out))
    ra = self.reduce_axis
    if ra is None:
      return _hp_sum(out) / norm
    if len(ra) == 0:
      return out / norm
    if not sparse:
      return _hp_sum(out, tuple(a for a in ra)) / norm
    if neighbor.format is partition.OrderedSparse:
      raise ValueError('Cannot report per-particle values with an OrderedSparse neighbor list '
                       '(smap.py:969-974).')
    per_atom = torch.zeros(N, dtype=torch.float64, device=R.device)
    per_atom.index_add_(0, recv[mask], out[mask].double())       # segment_sum on idx[0]
    return per_atom.to(R.dtype) / norm


def _hp_sum(x, axis=None):
  if axis is None:
    return x.sum(dtype=torch.float64).to(x.dtype)
  return x.sum(dim=axis, dtype=torch.float64).to(x.dtype)


def pair_neighbor_list(fn, displacement_or_metric, species=None,
                       reduce_axis=None, ignore_unused_parameters=False,
                       **kwargs):
  """smap.py:856-979.  For this package's pair potentials
  (`energy.lennard_jones`, `energy.soft_sphere`, `energy.morse`, optionally
  wrapped by `energy.multiplicative_isotropic_cutoff`) the mapped function is
  the fused CUDA kernel; any other Python `fn(dr, **params)` is mapped by
  `GenericPairNeighborListFn` (torch ops over the CUDA-built list)."""
  pot = getattr(fn, '_jmd_potential', None)
  if pot is None or any(_needs_generic(v) for v in kwargs.values()):
    return GenericPairNeighborListFn(fn, displacement_or_metric, species, reduce_axis,
                                     ignore_unused_parameters, kwargs)
  return PairNeighborListFn(pot, displacement_or_metric, species, reduce_axis,
                            ignore_unused_parameters, kwargs, fn=fn)
