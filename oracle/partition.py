"""Oracle: cell list + neighbour list (reference `jax_md/partition.py`).

Test infrastructure only.  A NumPy restatement that keeps the reference's op
order, slot layout, candidate order and capacity rules.  Citations are
`partition.py:line` in /root/reference/jax_md.
"""
import enum
from functools import reduce
from operator import mul

import numpy as np

from . import space

f32 = np.float32
i32 = np.int32


class PEC(enum.IntEnum):
  """partition.py:494-520."""
  NONE = 0
  NEIGHBOR_LIST_OVERFLOW = 1
  CELL_LIST_OVERFLOW = 2
  CELL_SIZE_TOO_SMALL = 4
  MALFORMED_BOX = 8


class Format(enum.Enum):
  """partition.py:641-657."""
  Dense = 0
  Sparse = 1
  OrderedSparse = 2


Dense, Sparse, OrderedSparse = Format.Dense, Format.Sparse, Format.OrderedSparse


def is_sparse(fmt):
  return fmt in (Sparse, OrderedSparse)


def err_update(code, bit, pred):
  """PartitionError.update, partition.py:535-541: OR `bit` in where pred."""
  return np.uint8(code | (np.uint8(bit) if bool(pred) else np.uint8(0)))


# ----------------------------------------------------------------------------
# Cell list
# ----------------------------------------------------------------------------

def cell_dimensions(dim, box_size, minimum_cell_size):
  """partition.py:146-188."""
  if isinstance(box_size, (int, float)):
    box_size = float(box_size)
  cells_per_side = np.floor(box_size / minimum_cell_size)
  cell_size = box_size / cells_per_side
  cells_per_side = np.array(cells_per_side, dtype=i32)
  if isinstance(box_size, np.ndarray):
    if box_size.ndim in (1, 2):
      assert box_size.size == dim
      flat = np.reshape(cells_per_side, (-1,))
      for c in flat:
        if c < 3:
          raise ValueError('Box must be at least 3x the size of the grid '
                           'spacing in each dimension.')
      cell_count = reduce(mul, [int(c) for c in flat], 1)
    elif box_size.ndim == 0:
      cell_count = int(cells_per_side) ** dim
    else:
      raise ValueError('bad box')
  else:
    cell_count = int(cells_per_side) ** dim
  return box_size, cell_size, cells_per_side, int(cell_count)


def hash_constants(dim, cells_per_side):
  """partition.py:212-224.  x is the fastest-varying cell coordinate."""
  if cells_per_side.size == 1:
    return np.array([[int(cells_per_side) ** d for d in range(dim)]], dtype=i32)
  cps = np.reshape(cells_per_side, (1, -1))
  one = np.array([[1]], dtype=i32)
  cps = np.concatenate((one, cps[:, :-1]), axis=1)
  return np.array(np.cumprod(cps), dtype=i32)[None, :]


def particle_cells(position, cell_size, cells_per_side):
  """partition.py:421-422: truncation toward zero, then non-negative mod."""
  idx = np.array(position / cell_size, dtype=i32)   # C cast == trunc
  return np.mod(idx, cells_per_side).astype(i32)


class CellList:
  """partition.py:78-133 (fields used by the neighbour list)."""

  def __init__(self, position_buffer, id_buffer, particle_cell_id, overflow,
               cell_capacity, cell_size, cells_per_side, max_occupancy):
    self.position_buffer = position_buffer   # [cell_count, cap, dim]
    self.id_buffer = id_buffer               # [cell_count, cap]
    self.particle_cell_id = particle_cell_id  # [N, dim]
    self.did_buffer_overflow = overflow
    self.cell_capacity = cell_capacity
    self.cell_size = cell_size
    self.cells_per_side = cells_per_side     # flat [dim]
    self.max_cell_occupancy = max_occupancy


def cell_list_build(position, box_size, minimum_cell_size,
                    buffer_size_multiplier=1.25, capacity=None,
                    extra_capacity=0):
  """partition.py:349-471 (`cell_list_fn`); `capacity=None` is `allocate`."""
  if isinstance(box_size, np.ndarray) and box_size.ndim == 1:
    box_size = np.reshape(box_size, (1, -1))                    # :333-336
  N, dim = position.shape
  if dim not in (2, 3):
    raise ValueError('Cell list spatial dimension must be 2 or 3.')
  _, cell_size, cells_per_side, cell_count = cell_dimensions(
      dim, box_size, minimum_cell_size)
  mult = hash_constants(dim, cells_per_side)
  indices = particle_cells(position, cell_size, cells_per_side)
  hashes = np.sum(indices * mult, axis=1).astype(i32)
  occupancy = np.bincount(hashes, minlength=cell_count)
  max_occ = int(occupancy.max()) if N else 0
  if capacity is None:                                           # :369-373
    capacity = int(max_occ * buffer_size_multiplier) + extra_capacity
  overflow = bool(max_occ > capacity)                            # :458-460

  sort_map = np.argsort(hashes, kind='stable')                   # :432
  sorted_hash = hashes[sort_map]
  slot = np.mod(np.arange(N, dtype=np.int64), max(capacity, 1))  # :441
  sorted_cell_id = sorted_hash.astype(np.int64) * capacity + slot
  cell_position = np.zeros((cell_count * capacity, dim), position.dtype)
  cell_id = N * np.ones((cell_count * capacity,), i32)
  if capacity > 0:
    # On overflow two atoms collide on one slot; XLA leaves the winner
    # unspecified, NumPy keeps the last write.  Only the flag is defined.
    cell_position[sorted_cell_id] = position[sort_map]
    cell_id[sorted_cell_id] = sort_map.astype(i32)
  flat_cps = np.broadcast_to(np.reshape(cells_per_side, (-1,)),
                             (dim,)).astype(i32) if cells_per_side.size == 1 \
      else np.reshape(cells_per_side, (-1,)).astype(i32)
  return CellList(cell_position.reshape(cell_count, capacity, dim),
                  cell_id.reshape(cell_count, capacity), indices, overflow,
                  capacity, cell_size, flat_cps, max_occ)


# ----------------------------------------------------------------------------
# Neighbour list
# ----------------------------------------------------------------------------

def cell_size_fn(box, minimum_cell_size):
  """partition.py:590-592."""
  cells_per_side = np.floor(box / minimum_cell_size)
  return box / cells_per_side


def fractional_cell_size(box, cutoff):
  """partition.py:595-638."""
  box = np.asarray(box, f32)
  cutoff = f32(cutoff)
  if box.ndim == 0:
    return cutoff / box
  if box.ndim == 1:
    return cutoff / np.min(box)
  if box.shape[0] == 1:
    return f32(1) / np.floor(box[0, 0] / cutoff)
  if box.shape[0] == 2:
    xx, yy = box[0, 0], box[1, 1]
    xy = box[0, 1] / yy
    nx, ny = xx / np.sqrt(f32(1) + xy ** 2), yy
    nmin = np.floor(np.min(np.array([nx, ny], f32)) / cutoff)
  else:
    xx, yy, zz = box[0, 0], box[1, 1], box[2, 2]
    xy, xz, yz = box[0, 1] / yy, box[0, 2] / zz, box[1, 2] / zz
    nx = xx / np.sqrt(f32(1) + xy ** 2 + (xy * yz - xz) ** 2)
    ny = yy / np.sqrt(f32(1) + yz ** 2)
    nmin = np.floor(np.min(np.array([nx, ny, zz], f32)) / cutoff)
  nmin = f32(1) if nmin == 0 else nmin
  return f32(1) / nmin


def is_box_valid(box):
  """partition.py:676-681."""
  box = np.asarray(box)
  if box.ndim in (0, 1):
    return True
  if box.ndim == 2:
    return bool(np.all(np.triu(box) == box))
  return False


def neighboring_cells(dim):
  """partition.py:232-240: first coordinate slowest, last fastest."""
  return np.array(list(np.ndindex(*([3] * dim))), dtype=i32) - 1


class NeighborList:
  """partition.py:684-737."""

  def __init__(self, idx, reference_position, error, cell_list_capacity,
               max_occupancy, format, cell_size, use_cell_list, owner):
    self.idx = idx
    self.reference_position = reference_position
    self.error = np.uint8(error)
    self.cell_list_capacity = cell_list_capacity
    self.max_occupancy = max_occupancy
    self.format = format
    self.cell_size = cell_size
    self.use_cell_list = use_cell_list
    self._owner = owner
    self.occupancy = None     # oracle-only diagnostics
    self.did_rebuild = True

  def update(self, position, **kw):
    return self._owner.update(position, self, **kw)

  @property
  def did_buffer_overflow(self):
    return np.uint8(self.error & (PEC.NEIGHBOR_LIST_OVERFLOW |
                                  PEC.CELL_LIST_OVERFLOW))


class neighbor_list:
  """partition.py:801-1164.  `displacement` is an oracle.space displacement
  function (positions-in, vector-out); metric^2 = sum(disp^2)."""

  def __init__(self, displacement, box, r_cutoff, dr_threshold=0.0,
               capacity_multiplier=1.25, disable_cell_list=False,
               mask_self=True, custom_mask_function=None,
               fractional_coordinates=False, format=Dense, chunk=4096,
               **static_kwargs):
    self.fractional = fractional_coordinates
    self.always_rebuild = (dr_threshold == 0)                    # :892
    self.box = f32(box) if np.ndim(box) == 0 else np.asarray(box, f32)  # :897
    self.cutoff = r_cutoff + dr_threshold                        # :899
    self.cutoff_sq = self.cutoff ** 2                            # :900
    self.threshold_sq = (dr_threshold / f32(2)) ** 2             # :901
    self.metric_sq = space.metric_sq(displacement)
    self._mkw = {}
    self.capacity_multiplier = capacity_multiplier
    self.disable_cell_list = disable_cell_list
    self.mask_self = mask_self
    self.custom_mask_function = custom_mask_function
    self.format = format
    self.chunk = chunk

  # -- candidates ------------------------------------------------------------
  def _cmp(self, d2, dtype):
    # `dR < cutoff_sq`: a Python-float cutoff is weakly typed -> compared in
    # the array dtype; an f32 cutoff against f64 d2 promotes to f64.
    c = self.cutoff_sq
    if isinstance(c, float):
      c = dtype.type(c)
    return d2 < c

  def _cell_candidates(self, cl, position, lo, hi):
    """partition.py:911-951 for atoms lo..hi-1 -> [n, 3^d*cap] ids (N=masked)."""
    N, dim = position.shape
    cps = cl.cells_per_side
    shifts = neighboring_cells(dim)
    shifted = cl.particle_cell_id[lo:hi, None, :] + shifts[None, :, :]
    shifted = np.mod(shifted, cps[None, None, :])
    mult = np.cumprod(np.concatenate(([1], cps[:-1]))).astype(np.int64)
    cell_1d = np.sum(shifted * mult, axis=2)                     # [n, 3^d]
    ids = cl.id_buffer[cell_1d].reshape(hi - lo, -1)
    pos = cl.position_buffer[cell_1d].reshape(hi - lo, -1, dim)
    d2 = self.metric_sq(position[lo:hi, None, :], pos, **self._mkw)
    return np.where(self._cmp(d2, position.dtype), ids, N).astype(i32)

  def _build(self, position, err, neighbors, extra_capacity, max_occupancy, **kwargs):
    """`neighbor_fn`, partition.py:1037-1117."""
    N, dim = position.shape
    cl = None
    cell_size = None
    use_cells = False
    self._mkw = {k: v for k, v in kwargs.items() if k == 'box'}   # metric kwargs (:1063)
    if not self.disable_cell_list:
      if neighbors is None:
        _box = kwargs.get('box', self.box)
        cell_size = self.cutoff
        if self.fractional:                                      # :1047-1051
          err = err_update(err, PEC.MALFORMED_BOX, is_box_valid(_box))
          cell_size = fractional_cell_size(_box, self.cutoff)
          _box = 1.0
        if np.all(np.asarray(cell_size) < _box / 3.0):           # :1052
          cl = cell_list_build(position, _box, cell_size,
                               self.capacity_multiplier, None, extra_capacity)
      else:
        cell_size = neighbors.cell_size
        if neighbors.use_cell_list:
          cl = cell_list_build(position, 1.0 if self.fractional else self.box, cell_size,
                               self.capacity_multiplier,
                               neighbors.cell_list_capacity)
      use_cells = cl is not None
    cl_capacity = None
    if use_cells:
      err = err_update(err, PEC.CELL_LIST_OVERFLOW, cl.did_buffer_overflow)
      cl_capacity = cl.cell_capacity
      W = 3 ** dim * cl.cell_capacity
    else:
      W = N

    fmt = self.format
    rows = np.arange(N, dtype=i32)
    dense_out = None
    recv, send = [], []
    occupancy = 0
    if not is_sparse(fmt):
      dense_out = np.full((N, W), N, i32)
    for lo in range(0, N, self.chunk):
      hi = min(N, lo + self.chunk)
      if use_cells:
        idx = self._cell_candidates(cl, position, lo, hi)
      else:
        idx = np.broadcast_to(np.arange(N, dtype=i32)[None, :],
                              (hi - lo, N)).copy()               # :904-909
      if self.mask_self:                                         # :953-958
        idx = np.where(idx == rows[lo:hi, None], N, idx)
      if self.custom_mask_function is not None:
        raise NotImplementedError('oracle: chunked custom mask')
      if is_sparse(fmt):
        sender = np.broadcast_to(rows[lo:hi, None], idx.shape)
        if use_cells:                                            # :1010-1032
          mask = idx < N
        else:                                                    # :982-1008
          jj = np.minimum(idx, N - 1)
          d2 = self.metric_sq(position[lo:hi, None, :], position[jj], **self._mkw)
          mask = self._cmp(d2, position.dtype) & (idx < N)
        if fmt is OrderedSparse:
          mask = mask & (idx < sender)
        recv.append(idx[mask])
        send.append(sender[mask])
        occupancy += int(mask.sum())
      else:                                                      # :960-980
        jj = np.minimum(idx, N - 1)           # OOB gather clamps
        # map_neighbor evaluates d(R_j, R_i) (space.py:494-502)
        d2 = self.metric_sq(position[jj], position[lo:hi, None, :], **self._mkw)
        mask = self._cmp(d2, position.dtype) & (idx < N)
        cs = np.cumsum(mask, axis=1)
        r, c = np.nonzero(mask)
        dense_out[lo + r, cs[r, c] - 1] = idx[r, c]
        occupancy = max(occupancy, int(cs[:, -1].max()) if W else 0)

    if is_sparse(fmt):
      width = N * W
      recv = np.concatenate(recv) if recv else np.zeros((0,), i32)
      send = np.concatenate(send) if send else np.zeros((0,), i32)
    else:
      width = W

    if max_occupancy is None:                                    # :1090-1104
      _extra = extra_capacity if not is_sparse(fmt) else N * extra_capacity
      max_occupancy = int(occupancy * self.capacity_multiplier + _extra)
      if max_occupancy > width:
        max_occupancy = width
      if not is_sparse(fmt):
        limit = N - 1 if self.mask_self else N
      elif fmt is Sparse:
        limit = N * (N - 1) if self.mask_self else N ** 2
      else:
        limit = N * (N - 1) // 2
      if max_occupancy > limit:
        max_occupancy = limit

    if is_sparse(fmt):
      idx = np.full((2, max_occupancy), N, i32)
      k = min(max_occupancy, occupancy)
      idx[0, :k] = recv[:k]
      idx[1, :k] = send[:k]
    else:
      idx = dense_out[:, :max_occupancy]                         # :1105
    err = err_update(err, PEC.NEIGHBOR_LIST_OVERFLOW,
                     occupancy > max_occupancy)                  # :1110
    nl = NeighborList(idx, position.copy(), err, cl_capacity, max_occupancy,
                      fmt, cell_size, use_cells, self)
    nl.occupancy = occupancy
    nl.max_cell_occupancy = cl.max_cell_occupancy if use_cells else None
    return nl

  # -- public ----------------------------------------------------------------
  def allocate(self, position, extra_capacity=0, **kw):
    """partition.py:1156-1157."""
    return self._build(position, np.uint8(0), None, extra_capacity, None, **kw)

  def needs_rebuild(self, position, neighbors, **kw):
    """partition.py:1146-1154 predicate (strict >)."""
    mkw = {k: v for k, v in kw.items() if k == 'box'}
    d2 = self.metric_sq(position, neighbors.reference_position, **mkw)
    t = self.threshold_sq
    if isinstance(t, float):
      t = position.dtype.type(t)
    return bool(np.any(d2 > t))

  def update(self, position, neighbors, **kw):
    """partition.py:1159-1160 / 1119-1154."""
    if 'box' in kw and not self.disable_cell_list:               # :1125-1139
      if not self.fractional:
        raise ValueError('Neighbor list cannot accept a box keyword argument if '
                         'fractional_coordinates is not enabled.')
      err = neighbors.error
      if neighbors.use_cell_list:
        cur = cell_size_fn(1.0, neighbors.cell_size)
        new = cell_size_fn(1.0, fractional_cell_size(kw['box'], self.cutoff))
        err = err_update(err, PEC.CELL_SIZE_TOO_SMALL, bool(np.any(new > cur)))
      err = err_update(err, PEC.MALFORMED_BOX, is_box_valid(kw['box']))
      neighbors.error = err
    if self.always_rebuild or self.needs_rebuild(position, neighbors, **kw):
      return self._build(position, neighbors.error, neighbors, 0,
                         neighbors.max_occupancy, **kw)
    neighbors.did_rebuild = False
    return neighbors


def neighbor_list_mask(neighbor, mask_self=False):
  """partition.py:1167-1182."""
  N = len(neighbor.reference_position)
  if is_sparse(neighbor.format):
    mask = neighbor.idx[0] < N
    if mask_self:
      mask = mask & (neighbor.idx[0] != neighbor.idx[1])
    return mask
  mask = neighbor.idx < len(neighbor.idx)
  if mask_self:
    mask = mask & (neighbor.idx != np.arange(N, dtype=i32)[:, None])
  return mask
