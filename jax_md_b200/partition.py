"""Neighbour lists: drop-in for the reference `jax_md/partition.py`
(`neighbor_list`, `NeighborList`, `NeighborListFns`, `NeighborListFormat`,
`PartitionError(Code)`, `neighbor_list_mask`, `is_sparse`).

Host logic (capacity rules, the cell-list decision, error-bit bookkeeping)
follows partition.py:801-1164 line by line; all array work is done by
libjmd_b200.so (csrc/jmd_neighbor.cu).  `allocate` syncs with the device to
read occupancies, as the reference does (partition.py:249,1094); `update` never
syncs and is CUDA-graph capturable.
"""
import copy as _copy
import ctypes as C
import logging
import os
from enum import Enum, IntEnum
from typing import Any, Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _lib, dataclasses, space

f32 = np.float32
i32 = np.int32
_FUSED_SKIN = os.environ.get('JMD_FUSED_SKIN', '1') != '0'    # debugging / A-B switch


class PartitionErrorCode(IntEnum):
  """partition.py:494-520."""
  NONE = 0
  NEIGHBOR_LIST_OVERFLOW = 1 << 0
  CELL_LIST_OVERFLOW = 1 << 1
  CELL_SIZE_TOO_SMALL = 1 << 2
  MALFORMED_BOX = 1 << 3


PEC = PartitionErrorCode


@dataclasses.dataclass
class PartitionError:
  """partition.py:523-561.  `code` is a uint8 device scalar."""
  code: Any

  def update(self, bit, pred):
    zero = torch.zeros_like(self.code)
    bit_t = torch.full_like(self.code, int(bit))
    pred_t = torch.as_tensor(pred, device=self.code.device)
    return PartitionError(self.code | torch.where(pred_t, bit_t, zero))

  def __str__(self):
    code = int(self.code)
    if code == PEC.NONE:
      return ''
    if code & PEC.NEIGHBOR_LIST_OVERFLOW:
      return 'Partition Error: Neighbor list buffer overflow.'
    if code & PEC.CELL_LIST_OVERFLOW:
      return 'Partition Error: Cell list buffer overflow'
    if code & PEC.CELL_SIZE_TOO_SMALL:
      return 'Partition Error: Cell size too small'
    if code & PEC.MALFORMED_BOX:
      return 'Partition Error: Incorrect box format. Expecting upper triangular.'
    raise ValueError(f'Unexpected error code {code}.')

  __repr__ = __str__


class NeighborListFormat(Enum):
  """partition.py:641-657."""
  Dense = 0
  Sparse = 1
  OrderedSparse = 2


Dense = NeighborListFormat.Dense
Sparse = NeighborListFormat.Sparse
OrderedSparse = NeighborListFormat.OrderedSparse


def is_sparse(fmt: NeighborListFormat) -> bool:
  return fmt is Sparse or fmt is OrderedSparse


def is_format_valid(fmt):
  if fmt not in list(NeighborListFormat):
    raise ValueError('Neighbor list format must be a member of '
                     f'NeighborListFormat found {fmt}.')


# ----------------------------------------------------------------------------
# device workspace: the hidden part of a NeighborList
# ----------------------------------------------------------------------------

class Workspace:
  """Owns the device buffers behind one allocated neighbour list and the
  `jmd_nbr_t` descriptor handed to the C ABI."""

  def __init__(self, n, dim, dtype, device):
    self.n, self.dim, self.dtype, self.device = n, dim, dtype, device
    self.c = _lib.NbrT()
    self.t = {}          # name -> tensor (keeps buffers alive)
    self.species = None
    self.update_mode = 'fused'   # 'fused' (one cooperative kernel) | 'gated'

  def buf(self, name, shape, dtype, fill=None):
    if fill is None:
      t = torch.empty(shape, dtype=dtype, device=self.device)
    else:
      t = torch.full(shape, fill, dtype=dtype, device=self.device)
    self.t[name] = t
    setattr(self.c, name, t.data_ptr())
    return t

  def buf_plain(self, name, shape, dtype):
    """Scratch owned by the workspace but not part of the C descriptor."""
    t = torch.empty(shape, dtype=dtype, device=self.device)
    self.t[name] = t
    return t

  def ref(self):
    return C.byref(self.c)

  def set_species(self, species):
    """Registers per-atom species ids (copied into pos_sorted.w on every
    refresh); None clears it."""
    if species is None:
      self.species = None
      self.c.species = None
      return
    if species.dtype != torch.int32 or not species.is_contiguous():
      species = species.to(torch.int32).contiguous()
    self.species = species
    self.c.species = species.data_ptr()

  def set_box(self, spec, box):
    """The box of ONE call on a periodic_general space: `box=` when the call passes it, else the
    box the displacement function was made with -- the reference's calls are functional, a
    `box=` of an earlier call never carries over (space.py:419-433, partition.py:1145)."""
    if not getattr(spec, 'general', False):
      return
    if box is None:
      if getattr(self, '_box_token', None) == 'default':
        return
      self._box_token = 'default'
      self.c.space = space.space_struct(spec, self.dim, self.dtype)
      return
    self._box_token = 'call'
    b = box.detach().cpu().numpy() if isinstance(box, torch.Tensor) else box
    self.c.space = space.space_struct(spec._replace(side=b), self.dim, self.dtype)

  def state_host(self):
    out = (C.c_int64 * _lib.ST_COUNT)()
    _lib.call('jmd_nbr_state_host', self.ref(), out, _lib.stream())
    return list(out)


@dataclasses.dataclass
class NeighborList:
  """partition.py:684-737 (same field names / static-dynamic split).

  `idx` is materialised in the requested format on every rebuild.  `_ws` is
  the hidden device workspace (cell-sorted float4 positions, permutation and
  the transposed full-row list the force kernels read); updates reuse these
  buffers in place, like the reference under jit with donated buffers."""
  idx: Any
  reference_position: Any
  error: PartitionError
  cell_list_capacity: Optional[int] = dataclasses.static_field()
  max_occupancy: int = dataclasses.static_field()
  format: NeighborListFormat = dataclasses.static_field()
  cell_size: Any = dataclasses.static_field()
  cell_list_fn: Any = dataclasses.static_field()
  update_fn: Callable = dataclasses.static_field()
  _ws: Any = dataclasses.static_field(default=None)

  def update(self, position, **kwargs) -> 'NeighborList':
    return self.update_fn(position, self, **kwargs)

  @property
  def did_buffer_overflow(self):
    return self.error.code & (PEC.NEIGHBOR_LIST_OVERFLOW | PEC.CELL_LIST_OVERFLOW)

  @property
  def cell_size_too_small(self):
    return self.error.code & PEC.CELL_SIZE_TOO_SMALL

  @property
  def malformed_box(self):
    return self.error.code & PEC.MALFORMED_BOX

  @property
  def internal_list_is_current(self) -> bool:
    """True while `idx` is the array our builder wrote (nobody swapped it)."""
    return self._ws is not None and self._ws.t.get('idx') is self.__dict__['idx']


def _idx_get(self):
  """`NeighborList.idx`.  With `lazy_idx=True` a rebuild inside update() only
  marks the public array stale on the device; reading the attribute enqueues the
  (device-gated) export first, so what the caller sees is always the array of
  the current list in the requested format."""
  v = self.__dict__['idx']
  ws = self.__dict__.get('_ws')
  if ws is not None and ws.c.lazy_idx and ws.t.get('idx') is v and not ws.c.no_public_idx:
    _lib.call('jmd_nbr_export', ws.ref(), None, 2, _lib.stream())
  return v


def _idx_set(self, value):
  self.__dict__['idx'] = value


# a data descriptor on the frozen dataclass: __init__ / replace() store through it
NeighborList.idx = property(_idx_get, _idx_set)


@dataclasses.dataclass
class NeighborListFns:
  """partition.py:740-785."""
  allocate: Callable = dataclasses.static_field()
  update: Callable = dataclasses.static_field()

  def __call__(self, position, neighbors=None, extra_capacity: int = 0,
               **kwargs):
    logging.warning('Using a deprecated code path to create / update neighbor '
                    'lists. Using `neighbor_fn.allocate` and '
                    '`neighbor_fn.update` is preferred.')
    if neighbors is None:
      return self.allocate(position, extra_capacity, **kwargs)
    return self.update(position, neighbors, **kwargs)

  def __iter__(self):
    return iter((self.allocate, self.update))


def _cell_dimensions(spatial_dimension, box_size, minimum_cell_size):
  """partition.py:146-188 (NumPy on the host, as in the reference)."""
  if isinstance(box_size, (int, float)):
    box_size = float(box_size)
  cells_per_side = np.floor(box_size / minimum_cell_size)
  cell_size = box_size / cells_per_side
  cells_per_side = np.array(cells_per_side, dtype=i32)
  if isinstance(box_size, np.ndarray):
    if box_size.ndim == 1 or box_size.ndim == 2:
      assert box_size.size == spatial_dimension
      flat = np.reshape(cells_per_side, (-1,))
      for cells in flat:
        if cells < 3:
          raise ValueError('Box must be at least 3x the size of the grid '
                           'spacing in each dimension.')
      cell_count = int(np.prod(flat.astype(np.int64)))
    elif box_size.ndim == 0:
      cell_count = int(cells_per_side) ** spatial_dimension
    else:
      raise ValueError('Box must be either: a scalar, a vector, or a matrix. '
                       f'Found {box_size}.')
  else:
    cell_count = int(cells_per_side) ** spatial_dimension
  return box_size, cell_size, cells_per_side, int(cell_count)


@dataclasses.dataclass
class CellList:
  """partition.py:79-131 (same fields).  Buffers have shape S = cells_per_side... +
  [cell_capacity]; empty slots hold id N."""
  position_buffer: Any
  id_buffer: Any
  particle_cell_id: Any
  named_buffer: Any
  did_buffer_overflow: Any
  cell_capacity: int = dataclasses.static_field()
  cell_size: Any = None
  update_fn: Callable = dataclasses.static_field(default=None)

  def update(self, position, **kwargs) -> 'CellList':
    cl_data = (self.cell_capacity, self.did_buffer_overflow, self.update_fn)
    return self.update_fn(position, cl_data, **kwargs)

  @property
  def kwarg_buffers(self):
    logging.warning('kwarg_buffers renamed to named_buffer. The name kwarg_buffers will be depricated.')
    return self.named_buffer


@dataclasses.dataclass
class CellListFns:
  """partition.py:134-143."""
  allocate: Callable = dataclasses.static_field()
  update: Callable = dataclasses.static_field()

  def __iter__(self):
    return iter((self.allocate, self.update))


def cell_list(box_size, minimum_cell_size, buffer_size_multiplier: float = 1.25) -> CellListFns:
  """The PUBLIC cell list of the reference (partition.py:296-488): dense per-cell buffers of
  positions, ids and per-particle side data (`**kwargs`), same shapes, slot rule
  (`sorted rank mod capacity`, :441), padding value and overflow flag.  It is a data-format
  utility next to the hot path -- `neighbor_list` never materialises these buffers, its
  kernels keep a CSR cell order instead -- so it is composed from device tensor ops
  (stable sort + scatter, like the reference's argsort + `.at[].set`)."""
  box_np = _host_scalar(box_size)
  if isinstance(box_np, np.ndarray) and box_np.ndim == 1:
    box_np = np.reshape(box_np, (1, -1))
  min_cell = _host_scalar(minimum_cell_size)

  def cell_list_fn(position, capacity_overflow_update=None, extra_capacity=0, **kwargs):
    _lib.require_cuda()
    N, dim = position.shape
    if dim not in (2, 3):
      raise ValueError(f'Cell list spatial dimension must be 2 or 3. Found {dim}.')
    _, cell_size, cells_per_side, cell_count = _cell_dimensions(dim, box_np, min_cell)
    cps = np.broadcast_to(np.reshape(cells_per_side, (-1,)), (dim,)).astype(np.int64)
    dev = position.device
    cs = torch.as_tensor(np.broadcast_to(np.reshape(np.asarray(cell_size, np.float64), (-1,)), (dim,)).copy(),
                         dtype=position.dtype, device=dev)
    cps_t = torch.as_tensor(cps, device=dev)
    indices = torch.remainder((position / cs).to(torch.int32), cps_t.to(torch.int32))     # :421-422
    mult = torch.as_tensor(np.concatenate([[1], np.cumprod(cps[:-1])]), device=dev)       # x fastest (:212-224)
    hashes = (indices.long() * mult).sum(1)
    if capacity_overflow_update is None:
      occ = torch.bincount(hashes, minlength=cell_count)
      cell_capacity = int(int(occ.max()) * buffer_size_multiplier) + extra_capacity       # :243-250, 369-373
      overflow = torch.zeros((), dtype=torch.bool, device=dev)
      update = cell_list_fn
    else:
      cell_capacity, overflow, update = capacity_overflow_update
      overflow = torch.as_tensor(overflow, device=dev, dtype=torch.bool)
    sort_map = torch.sort(hashes, stable=True).indices                                    # :432
    sorted_cell_id = hashes[sort_map] * cell_capacity + \
        torch.remainder(torch.arange(N, device=dev), cell_capacity)                       # :441-442
    shape = tuple(int(c) for c in cps[::-1]) + (cell_capacity,)
    pos_buf = torch.zeros((cell_count * cell_capacity, dim), dtype=position.dtype, device=dev)
    id_buf = torch.full((cell_count * cell_capacity, 1), N, dtype=torch.int32, device=dev)
    pos_buf[sorted_cell_id] = position[sort_map]
    id_buf[sorted_cell_id] = sort_map.to(torch.int32)[:, None]
    named = {}
    for k, v in kwargs.items():
      if not isinstance(v, torch.Tensor):
        raise ValueError(f'Data must be specified as an ndarray. Found "{k}" with type {type(v)}.')
      if v.shape[0] != N:
        raise ValueError(f'Data must be specified per-particle (an ndarray with shape ({N}, ...)). '
                         f'Found "{k}" with shape {tuple(v.shape)}.')
      tail = tuple(v.shape[1:]) if v.ndim > 1 else (1,)
      buf = torch.full((cell_count * cell_capacity,) + tail, 10 ** 5, dtype=v.dtype, device=dev)
      buf[sorted_cell_id] = v[sort_map].reshape((N,) + tail)
      named[k] = buf.reshape(shape + tail)
    occ = torch.bincount(hashes, minlength=cell_count)
    overflow = overflow | (occ.max() > cell_capacity)                                     # :458-460
    return CellList(pos_buf.reshape(shape + (dim,)), id_buf.reshape(shape + (1,)), indices, named,
                    overflow, cell_capacity, cell_size, update)

  def allocate_fn(position, extra_capacity: int = 0, **kwargs):
    return cell_list_fn(position, extra_capacity=extra_capacity, **kwargs)

  def update_fn(position, cl_or_capacity, **kwargs):
    if isinstance(cl_or_capacity, int):
      return cell_list_fn(position, (int(cl_or_capacity), False, cell_list_fn), **kwargs)
    cl = cl_or_capacity
    return cell_list_fn(position, (cl.cell_capacity, cl.did_buffer_overflow, cl.update_fn), **kwargs)

  return CellListFns(allocate_fn, update_fn)


def _cell_size(box, minimum_cell_size):
  """partition.py:590-592."""
  cells_per_side = np.floor(box / minimum_cell_size)
  return box / cells_per_side


def _fractional_cell_size(box, cutoff):
  """partition.py:595-638 (f32 host arithmetic)."""
  box = np.asarray(box, f32) if not np.isscalar(box) else f32(box)
  cutoff = f32(cutoff)
  if np.ndim(box) == 0:
    return cutoff / box
  if box.ndim == 1:
    return cutoff / np.min(box)
  if box.ndim == 2:
    if box.shape[0] == 1:
      return f32(1) / np.floor(box[0, 0] / cutoff)
    if box.shape[0] == 2:
      xx, yy = box[0, 0], box[1, 1]
      xy = box[0, 1] / yy
      nx, ny = xx / np.sqrt(f32(1) + xy ** 2), yy
      nmin = np.floor(np.min(np.array([nx, ny], f32)) / cutoff)
      nmin = f32(1) if nmin == 0 else nmin
      return f32(1) / nmin
    if box.shape[0] == 3:
      xx, yy, zz = box[0, 0], box[1, 1], box[2, 2]
      xy, xz, yz = box[0, 1] / yy, box[0, 2] / zz, box[1, 2] / zz
      nx = xx / np.sqrt(f32(1) + xy ** 2 + (xy * yz - xz) ** 2)
      ny = yy / np.sqrt(f32(1) + yz ** 2)
      nz = zz
      nmin = np.floor(np.min(np.array([nx, ny, nz], f32)) / cutoff)
      nmin = f32(1) if nmin == 0 else nmin
      return f32(1) / nmin
    raise ValueError(f'Expected box to be either 1-, 2-, or 3-dimensional found {box.shape[0]}')
  raise ValueError(f'Expected box to be either a scalar, a vector, or a matrix. Found {type(box)}.')


def is_box_valid(box) -> bool:
  """partition.py:676-681."""
  box = np.asarray(box)
  if box.ndim in (0, 1):
    return True
  if box.ndim == 2:
    return bool(np.all(np.triu(box) == box))
  return False


def _host_scalar(x):
  """Python / NumPy scalar view of a (possibly torch) scalar, keeping dtype."""
  if isinstance(x, torch.Tensor):
    return x.detach().cpu().numpy()[()]
  return x


def workspace_buffers(c, n_buf, dim, n_cells):
  """The device buffers behind a neighbour list as (name, shape, kind, fill) with kind in
  'i4' | 'i8' | 'u1' | 'f' (position dtype); fill None = uninitialised.  Framework
  agnostic: the torch host (`Workspace`) and the XLA-FFI binding (`_jax_binding.py`)
  allocate from this one table.  `nl` and `idx` are sized after the occupancy pass."""
  n_cells_buf = c.n_fine_cells
  return [
      ('cell_count', (n_cells_buf + 1,), 'i4', 0),
      ('cell_start', (n_cells_buf + 1,), 'i4', 0),
      ('cell_cursor', (max(n_cells_buf, 1),), 'i4', 0),
      ('ref_count', (max(n_cells, 1),), 'i4', 0),
      ('ref_start', (max(n_cells, 1) + 1,), 'i4', 0),
      # two int scans or one int64 scan
      ('scan_tmp', (2 * (max(n_cells_buf, n_buf) // 2048 + 2) + 16,), 'i4', 0),
      ('hash', (max(n_buf, 1),), 'i4', None),
      ('tmp_ids', (max(n_buf, 1),), 'i4', None),
      ('perm', (c.n_pad,), 'i4', 0),
      ('inv_perm', (max(n_buf, 1),), 'i4', None),
      ('pos_sorted', (c.n_pad, 4), 'f', 0),
      ('cnt', (c.n_pad,), 'i4', 0),
      ('cnt_lower', (c.n_pad,), 'i4', 0),
      ('offsets', (n_buf + 1,), 'i8', 0),
      ('reference_position', (n_buf, dim), 'f', None),
      ('error', (), 'u1', 0),
      ('state', (_lib.ST_COUNT,), 'i8', 0),
      # skin predicate fused into the drift kernel (simulate._Stepper.step)
      ('skin_blk', (c.n_pad // 256 + 1,), 'i4', 0),
      # look-back words of the sparse offsets scan (csrc/jmd_nbr_cellscan.cuh)
      ('cs_lb', (n_buf // 2048 + 4,), 'i8', 0),
  ]


def capacity_rule(format, N, width, mask_self, capacity_multiplier, extra_capacity, max_row, total):
  """partition.py:1090-1104 -> (max_occupancy, m_int): the public capacity and the
  internal row capacity (Dense: the same; sparse formats: the Dense-equivalent rule on
  the longest row, DESIGN.md "row capacity")."""
  sparse = is_sparse(format)
  occupancy = total if sparse else max_row
  full_width = N * width if sparse else width
  _extra = extra_capacity if not sparse else N * extra_capacity
  max_occupancy = int(occupancy * capacity_multiplier + _extra)
  if max_occupancy > full_width:
    max_occupancy = full_width
  if not sparse:
    capacity_limit = N - 1 if mask_self else N
  elif format is Sparse:
    capacity_limit = N * (N - 1) if mask_self else N ** 2
  else:
    capacity_limit = N * (N - 1) // 2
  if max_occupancy > capacity_limit:
    max_occupancy = capacity_limit
  if sparse:
    m_int = min(int(max_row * capacity_multiplier + extra_capacity), width,
                N - 1 if mask_self else N)
  else:
    m_int = max_occupancy
  return max_occupancy, max(m_int, 1)


_CELL_SCAN = os.environ.get('JMD_CELL_SCAN', '0') != '0'
_CELL_SCAN_MIN_OCCUPANCY = float(os.environ.get('JMD_CELL_SCAN_MIN_OCC', '8'))


def _enable_cell_scan(ws, cl_capacity, static_kwargs):
  """Warp-per-cell candidate scan (csrc/jmd_nbr_cellscan.cuh): one warp tests the
  concatenated candidate stream of a home cell's 3^d stencil, 32 candidates at a
  time.  It pays when a cell holds enough atoms to fill the lanes (LJ liquid: ~20
  per cell); sparse cells (2-D soft spheres: ~2) keep the thread-per-atom scan.
  `cell_scan=True/False` (static kwarg) overrides the default, which is OFF: measured on
  B200 (LJ, N=1M, profiles/r02_*) the test loop needs 0.47 warp instructions per candidate
  against 1.1 for the thread-per-atom scan (0.45 vs 0.78 ms), but turning the accept masks
  into rows costs more than it saves (0.81-0.89 ms with the rows; rebuild 1.31 vs 1.21 ms)."""
  c = ws.c
  c.cell_scan = 0
  stage = static_kwargs.get('stage_positions', os.environ.get('JMD_STAGE', '0') != '0')
  if not ws.use_cells or ws.fine != 1 or c.brick_shift != 0 or stage:
    return
  want = static_kwargs.get('cell_scan')
  if want is None:
    want = _CELL_SCAN and ws.n_capacity / max(ws.n_cells_ref, 1) >= _CELL_SCAN_MIN_OCCUPANCY
  if not want:
    return
  c.cs_batches = -(-max(cl_capacity, 1) // 32)
  c.cs_chunks = -(-(3 ** ws.dim) * max(cl_capacity, 1) // 32)
  c.cell_scan = 1


def neighbor_list(displacement_or_metric,
                  box,
                  r_cutoff,
                  dr_threshold=0.0,
                  capacity_multiplier: float = 1.25,
                  disable_cell_list: bool = False,
                  mask_self: bool = True,
                  custom_mask_function=None,
                  fractional_coordinates: bool = False,
                  format: NeighborListFormat = NeighborListFormat.Dense,
                  **static_kwargs) -> NeighborListFns:
  """partition.py:801-1164.  Same arguments, same allocate/update contract,
  same capacity rules and error bits; the displacement function must come from
  `jax_md_b200.space` so its metric can be inlined into the kernels."""
  is_format_valid(format)
  if custom_mask_function is not None:
    return _masked_neighbor_list(displacement_or_metric, box, r_cutoff, dr_threshold,
                                 capacity_multiplier, disable_cell_list, mask_self,
                                 custom_mask_function, format, static_kwargs)
  spec = space.get_spec(displacement_or_metric)
  r_cutoff = _host_scalar(r_cutoff)
  dr_threshold = _host_scalar(dr_threshold)
  _always_rebuild = bool(dr_threshold == 0)                       # :892
  box_np = _host_scalar(box)
  box_np = f32(box_np) if np.ndim(box_np) == 0 else np.asarray(box_np, f32)  # :897
  if fractional_coordinates and not (spec.general and spec.fractional):
    raise ValueError('fractional_coordinates=True needs a displacement function from '
                     'space.periodic_general(box, fractional_coordinates=True)')
  # (a MATRIX box without fractional_coordinates=True stays a matrix here: the reference's test
  #  `all(cell_size < box / 3)`, partition.py:1052, then sees its zero off-diagonal elements and
  #  never builds a cell list -- all-pairs candidates, as in tests/simulate_test.py:591-620)
  # the box currently in force (fractional coordinates: `box=` overrides, partition.py:1045)
  current = {'box': box_np, 'metric_box': None}

  def _box_kwarg(b):
    """`box=` of allocate / update: the cell grid is sized from it (partition.py:1045) and a
    periodic_general metric uses it (`partial(metric_sq, **kwargs)`, :1145); other metrics
    ignore the keyword."""
    b = _host_scalar(b)
    if spec.general:
      current['metric_box'] = b                 # in the caller's precision, like the reference's metric
    b = f32(b) if np.ndim(b) == 0 else np.asarray(b, f32)
    current['box'] = b
    return b
  cutoff = r_cutoff + dr_threshold                                # :899
  cutoff_sq = cutoff ** 2                                         # :900
  threshold_sq = (dr_threshold / f32(2)) ** 2                     # :901
  fmt_code = {Dense: _lib.DENSE, Sparse: _lib.SPARSE,
              OrderedSparse: _lib.ORDERED_SPARSE}[format]

  def _typed(x, np_dtype):
    # `arr < python_float` compares in the array dtype (weak type); an f32
    # NumPy scalar against an f64 array promotes exactly.
    return float(np_dtype(x)) if isinstance(x, (float, int)) else float(x)

  def fill_descriptor(c, N, dim, np_dtype, n_buf):
    """Static half of a jmd_nbr_t for N atoms (host arithmetic only: the cell-list
    decision and grid of partition.py:1046-1054, metric constants, search-grid options).
    Shared by the torch host below and the XLA-FFI binding (_jax_binding.py).
    -> (use_cells, cell_size, n_reference_cells, fine)."""
    if dim not in (2, 3):
      raise ValueError(f'Cell list spatial dimension must be 2 or 3. Found {dim}.')
    c.n, c.dtype, c.format = N, (_lib.F32 if np_dtype == np.float32 else _lib.F64), fmt_code
    c.mask_self = 1 if mask_self else 0
    c.always_rebuild = 1 if _always_rebuild else 0
    c.cutoff_sq = _typed(cutoff_sq, np_dtype)
    c.threshold_sq = _typed(threshold_sq, np_dtype)
    _spec = spec if current['metric_box'] is None else spec._replace(side=current['metric_box'])
    c.space = space.space_struct(_spec, dim, torch.float32 if np_dtype == np.float32 else torch.float64)
    c.n_pad = ((n_buf + 31) // 32) * 32 if n_buf else 32

    use_cells, cell_size, cps, n_cells = False, None, np.ones(3, i32), 0
    if not disable_cell_list:
      cell_size = cutoff                                           # :1046
      _box = current['box']                                        # :1045 kwargs.get('box', box)
      if fractional_coordinates:                                   # :1047-1051
        cell_size = _fractional_cell_size(current['box'], cutoff)
        _box = 1.0
      if np.all(np.asarray(cell_size) < _box / 3.0):               # :1052
        if np.ndim(_box) == 2:
          raise NotImplementedError('cell grid for a matrix box in real-space coordinates: use '
                                    'fractional_coordinates=True (space.py:360-372)')
        _, cs, cpside, n_cells = _cell_dimensions(dim, _box, cell_size)
        use_cells = True
        cps[:dim] = np.broadcast_to(np.reshape(cpside, (-1,)), (dim,))
        cs = np.broadcast_to(np.reshape(np.asarray(cs, f32), (-1,)), (dim,))
        for k in range(dim):
          c.cell_size[k] = float(cs[k])
    c.use_cells = 1 if use_cells else 0
    for k in range(3):
      c.cps[k] = int(cps[k])
    c.n_cells = n_cells
    # Internal search grid: reference cells split in two per side, stencil +-2
    # fine cells (needs 2 * fine_size >= cutoff with a rounding margin and at
    # least 5 fine cells per side; otherwise fall back to the reference grid).
    # Measured on B200 (LJ, N=1M): the finer grid loses -- 2.5 atoms per fine
    # cell make the lanes of a warp diverge (rebuild 2.38 ms vs 1.96 ms) and the
    # fine-cell order costs the force kernel L1 locality (0.34 vs 0.29 ms) -- so
    # the default keeps the reference grid, which also keeps the reference's
    # candidate ORDER in `idx`.  `fine_search_grid=True` (static kwarg) opts in.
    fine, w = 1, 1
    if use_cells and static_kwargs.get('fine_search_grid', False):
      cs_min = min(c.cell_size[k] for k in range(dim))
      if cs_min >= float(cutoff) * (1.0 + 1e-4) and all(2 * cps[k] >= 5 for k in range(dim)):
        fine, w = 2, 2
    # Storage order of the cells (csrc/jmd_neighbor.cu "cell storage order"):
    # `cell_brick_shift=b` (static kwarg) stores cells in bricks of (2^b)^dim so
    # warps / blocks are spatially compact.  Measured on B200 (LJ, N=1M): no gain
    # (force kernel 0.27 ms vs 0.26 ms in the reference's x-fastest order, which
    # keeps the three x-neighbour cells of a stencil row contiguous), so the
    # default is the reference order; public `idx` is identical either way.
    bshift = int(static_kwargs.get('cell_brick_shift', os.environ.get('JMD_BRICK_SHIFT', 0))) if use_cells else 0
    brick = 1 << bshift
    n_fine = 1
    for k in range(3):
      c.fine_cps[k] = int(cps[k]) * fine if k < dim else 1
      if k < dim:
        n_fine *= -(-c.fine_cps[k] // brick) * brick
    for k in range(dim):
      c.fine_cell_size[k] = c.cell_size[k] / fine        # exact (power of two)
    c.n_fine_cells = n_fine if use_cells else 0
    c.brick_shift = bshift
    c.stencil_w = w
    # exact_scan=True (static kwarg): evaluate the reference arithmetic on every
    # candidate instead of only inside the pre-filter's rounding band (testing).
    c.no_filter = 1 if static_kwargs.get('exact_scan', False) else 0
    if spec.general and spec.fractional and not fractional_coordinates:
      # unit-cube positions binned on a REAL-space grid (what the reference does when the list is
      # made without fractional_coordinates=True, e.g. tests/simulate_test.py:593-620): every atom
      # lands in the corner cells, so a cell says nothing about where its atoms are and the
      # pre-filter's image shift does not apply: exact metric for every candidate
      c.no_filter = 1
    # lazy_idx=True (static kwarg): see NeighborList.idx
    c.lazy_idx = 1 if static_kwargs.get('lazy_idx', False) else 0
    return use_cells, cell_size, n_cells, fine

  def _make_workspace(position, extra_capacity, n_capacity=None):
    _lib.require_cuda()
    if not isinstance(position, torch.Tensor) or not position.is_cuda:
      raise TypeError('positions must be a CUDA torch.Tensor [N, dim]')
    N, dim = position.shape
    np_dtype = np.float32 if position.dtype == torch.float32 else np.float64
    # per-atom buffers can be over-allocated (domain decomposition: the local
    # atom count changes at every rebuild; see jax_md_b200/domain.py)
    n_buf = max(N, int(n_capacity or 0))
    ws = Workspace(N, dim, position.dtype, position.device)
    ws.n_capacity = n_buf
    c = ws.c
    use_cells, cell_size, n_cells, fine = fill_descriptor(c, N, dim, np_dtype, n_buf)
    kinds = {'i4': torch.int32, 'i8': torch.int64, 'u1': torch.uint8, 'f': position.dtype}
    for name, shape, kind, fill in workspace_buffers(c, n_buf, dim, n_cells):
      ws.buf(name, shape, kinds[kind], fill)
    ws.drift_out = None            # (weakref to the drift's position tensor, its version, update epoch)
    ws.update_epoch = 0            # bumped by every update(): drift flags older than that are stale
    ws.n_cells_ref = n_cells
    ws.fine = fine
    ws.cell_size_host = cell_size
    ws.use_cells = use_cells
    return ws

  def allocate_fn(position, extra_capacity: int = 0, n_capacity=None,
                  n_rows=None, no_public_idx=False, **kwargs):
    """partition.py:1156-1157 -> neighbor_fn with neighbors=None (not jittable:
    reads occupancies back to the host)."""
    if 'box' in kwargs:
      _box_kwarg(kwargs['box'])
    else:                                    # (a box= of an earlier call does not carry over)
      current['box'], current['metric_box'] = box_np, None
    position = position.contiguous()
    ws = _make_workspace(position, extra_capacity, n_capacity)
    ws._box_token = 'call' if 'box' in kwargs else 'default'
    if fractional_coordinates and not disable_cell_list and is_box_valid(current['box']):
      # partition.py:1049: `err.update(MALFORMED_BOX, is_box_valid(box))` -- the bit is set
      # when the box IS valid (reference quirk; replicated, not fixed)
      ws.t['error'].fill_(int(PEC.MALFORMED_BOX))
    c, N, dim = ws.c, ws.n, ws.dim
    c.n_rows = int(n_rows) if n_rows else 0
    c.no_public_idx = 1 if no_public_idx else 0
    st, pp = _lib.stream(), _lib.ptr(position)
    # -- cell capacity (partition.py:243-250, 369-373)
    c.cell_capacity = 1
    c.m_int, c.max_occupancy = 1, 1
    _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, st)
    cl_capacity = None
    if ws.use_cells:
      max_cell = ws.state_host()[_lib.ST_MAX_CELL_OCC]
      cl_capacity = int(max_cell * capacity_multiplier) + extra_capacity
      c.cell_capacity = cl_capacity
      width = 3 ** dim * cl_capacity
      _enable_cell_scan(ws, cl_capacity, static_kwargs)
      # cells are stored in the reference's slot order, which depends on the
      # capacity (slot = rank mod capacity, partition.py:441): bin again.
      _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, st)
    else:
      width = N
    # -- occupancy pass (partition.py:1083-1088)
    _lib.call('jmd_nbr_build', ws.ref(), pp, 1, 0, st)
    state = ws.state_host()
    max_row, total = state[_lib.ST_MAX_ROW], state[_lib.ST_TOTAL]
    sparse = is_sparse(format)
    max_occupancy, m_int = capacity_rule(format, N, width, mask_self, capacity_multiplier,
                                         extra_capacity, max_row, total)
    c.max_occupancy = max_occupancy
    c.m_int = max(m_int, 1)
    ws.buf('nl', (c.m_int, c.n_pad), torch.int32)
    # Force-kernel staging plan (csrc/jmd_common.cuh): 16-bit copy of the rows as
    # indices into the owning block's shared-memory staging buffer + the
    # per-block range table.  Opt-in (`stage_positions=True` static kwarg): measured
    # on B200 (LJ, N=1M) the staged kernel is not faster than the L1 gather kernel
    # (0.29-0.31 ms vs 0.26 ms) and emitting the 16-bit rows slows the rebuild.
    stage = static_kwargs.get('stage_positions', os.environ.get('JMD_STAGE', '0') != '0')
    if stage and ws.use_cells and c.stencil_w == 1 and c.brick_shift == 0:
      ws.buf('nl16', (c.m_int, c.n_pad), torch.uint16)
      ws.buf('blk_table', (c.n_pad // 256 + 1, 256), torch.int32, 0)
      c.staged = 1
    else:
      c.staged = 0
    if no_public_idx:
      idx = ws.buf('idx', (0,), torch.int32)
    elif sparse:
      idx = ws.buf('idx', (2, max_occupancy), torch.int32, N)
    else:
      idx = ws.buf('idx', (N, max_occupancy), torch.int32, N)
    if idx.numel() == 0:
      c.idx = None
    _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 0, st)
    _lib.call('jmd_nbr_export', ws.ref(), pp, 0, st)
    ref = ws.t['reference_position']
    return NeighborList(idx, ref if n_capacity is None else ref[:N],
                        PartitionError(ws.t['error']), cl_capacity,
                        max_occupancy, format, ws.cell_size_host,
                        'cell_list' if ws.use_cells else None, update_fn, ws)

  def update_fn(position, neighbors, **kwargs):
    """partition.py:1159-1160 / 1119-1154.  Never syncs with the host."""
    ws = neighbors._ws
    if 'box' in kwargs and not disable_cell_list:                  # partition.py:1125-1139
      if not fractional_coordinates:
        raise ValueError('Neighbor list cannot accept a box keyword argument if '
                         'fractional_coordinates is not enabled.')
      b = _box_kwarg(kwargs['box'])
      bits = 0
      if neighbors.cell_list_fn is not None:
        cur = _cell_size(1.0, neighbors.cell_size)
        new = _cell_size(1.0, _fractional_cell_size(b, cutoff))
        if np.any(new > cur):
          bits |= int(PEC.CELL_SIZE_TOO_SMALL)
      if is_box_valid(b):
        bits |= int(PEC.MALFORMED_BOX)
      if bits and ws is not None:
        ws.t['error'].bitwise_or_(torch.tensor(bits, dtype=torch.uint8, device=ws.t['error'].device))
      if ws is not None:
        # the metric of every later kernel (skin predicate, candidate tests, forces) uses the new box
        ws.c.space = space.space_struct(spec._replace(side=current['metric_box']), ws.dim, ws.dtype)
    elif 'box' in kwargs and ws is not None and spec.general:      # all-pairs: box only feeds the metric
      _box_kwarg(kwargs['box'])
      ws.c.space = space.space_struct(spec._replace(side=current['metric_box']), ws.dim, ws.dtype)
    if ws is not None and spec.general:
      if 'box' in kwargs:
        ws._box_token = 'call'
      else:
        ws.set_box(spec, None)          # no box= on this call: the displacement function's own box
    if ws is None:
      raise ValueError('This NeighborList was not allocated by jax_md_b200.')
    if position.shape != (ws.n, ws.dim) or position.dtype != ws.dtype:
      raise ValueError('position shape/dtype differs from the allocated list')
    position = position.contiguous()
    st, pp = _lib.stream(), _lib.ptr(position)
    # Did the drift kernel already evaluate the predicate for THIS tensor (same
    # object, not modified since) against THIS list's current reference positions
    # (no other update() ran in between)?  Then update() skips its own pass.
    tag, ws.drift_out = ws.drift_out, None
    fused_skin = (_FUSED_SKIN and tag is not None and tag[0]() is position
                  and tag[1] == position._version and tag[2] == ws.update_epoch)
    ws.update_epoch += 1
    if ws.update_mode == 'fused':
      ws.c.skin_pre = 1 if fused_skin else 0
      _lib.call('jmd_nbr_update', ws.ref(), pp, st)
      ws.c.skin_pre = 0
    else:
      _lib.call('jmd_nbr_skin_check', ws.ref(), pp, st)
      _lib.call('jmd_nbr_bin', ws.ref(), pp, 1, st)
      _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 1, st)
      _lib.call('jmd_nbr_export', ws.ref(), pp, 1, st)
    # partition.py:1107-1117 returns a new NeighborList; the buffers behind it are
    # this list's own, rewritten in place (like jit with donated arguments)
    return _copy.copy(neighbors)       # (not replace(): reading .idx would materialise a lazy idx)

  allocate_fn.fill_descriptor = fill_descriptor      # host planning shared with _jax_binding.py
  return NeighborListFns(allocate_fn, update_fn)


def _masked_neighbor_list(displacement_or_metric, box, r_cutoff, dr_threshold,
                          capacity_multiplier, disable_cell_list, mask_self,
                          custom_mask_function, format, static_kwargs):
  """`custom_mask_function` (partition.py:809, 1079-1080; tests/partition_test.py:403-459).

  The reference applies the user's function to the UNCOMPACTED candidate array and prunes
  by distance afterwards.  For the functions this hook exists for -- an entry is replaced
  by N depending on the (row, entry) pair, e.g. bonded-pair exclusions -- masking commutes
  with the distance pruning, so it is applied here to the kernel-built Dense rows, followed
  by the reference's stable compaction and capacity rule.  The result is a list whose `idx`
  is not the kernels' internal one: energies over it take the generic (torch-composed)
  pair path.  `update` re-masks only after the inner list rebuilt (one host read of the
  build counter, like the overflow check of the reference loop)."""
  inner = neighbor_list(displacement_or_metric, box, r_cutoff, dr_threshold, capacity_multiplier,
                        disable_cell_list, mask_self, None, False, Dense, **static_kwargs)
  sparse = is_sparse(format)

  def _from_dense(nb_inner, max_occupancy, cl_capacity):
    dense = nb_inner.idx
    N = dense.shape[0]
    masked = custom_mask_function(dense)
    # stable compaction of every row (partition.py:960-980): valid entries first
    order = torch.sort((masked >= N).to(torch.int8), dim=1, stable=True).indices
    rows = torch.gather(masked, 1, order)
    valid = rows < N
    if format is OrderedSparse:
      senders = torch.arange(N, device=rows.device, dtype=rows.dtype)[:, None]
      valid = valid & (rows < senders)                       # partition.py:1021-1022
    if sparse:
      occupancy = int(valid.sum())
    else:
      occupancy = int(valid.sum(1).max()) if N else 0
    if max_occupancy is None:
      w = N if cl_capacity is None else 3 ** nb_inner._ws.dim * cl_capacity
      max_occupancy, _ = capacity_rule(format, N, w, mask_self, capacity_multiplier,
                                       _masked_neighbor_list.extra, 0 if sparse else occupancy,
                                       occupancy)
    if sparse:
      senders = torch.arange(N, device=rows.device, dtype=torch.int32)[:, None].expand_as(rows)
      recv, send = rows[valid], senders[valid]
      idx = torch.full((2, max_occupancy), N, dtype=torch.int32, device=rows.device)
      k = min(max_occupancy, recv.numel())
      idx[0, :k] = recv[:k]
      idx[1, :k] = send[:k]
    else:
      idx = torch.full((N, max_occupancy), N, dtype=torch.int32, device=rows.device)
      k = min(max_occupancy, rows.shape[1])
      idx[:, :k] = torch.where(valid, rows, torch.full_like(rows, N))[:, :k]
    overflow = occupancy > max_occupancy
    return idx, max_occupancy, overflow

  def _wrap(nb_inner, idx, max_occupancy, overflow, builds):
    err = nb_inner.error.update(PEC.NEIGHBOR_LIST_OVERFLOW, overflow)
    out = NeighborList(idx, nb_inner.reference_position, err, nb_inner.cell_list_capacity,
                       max_occupancy, format, nb_inner.cell_size, nb_inner.cell_list_fn, update_fn, None)
    object.__setattr__(out, '_inner', nb_inner)
    object.__setattr__(out, '_inner_builds', builds)
    return out

  def allocate_fn(position, extra_capacity: int = 0, **kwargs):
    _masked_neighbor_list.extra = extra_capacity
    nb_inner = inner.allocate(position, extra_capacity, **kwargs)
    idx, max_occupancy, overflow = _from_dense(nb_inner, None, nb_inner.cell_list_capacity)
    return _wrap(nb_inner, idx, max_occupancy, overflow, nb_inner._ws.state_host()[_lib.ST_BUILDS])

  def update_fn(position, neighbors, **kwargs):
    nb_inner = neighbors._inner.update(position, **kwargs)
    builds = nb_inner._ws.state_host()[_lib.ST_BUILDS]
    if builds == neighbors._inner_builds:
      return neighbors                                         # partition.py:1146: identity branch
    idx, max_occupancy, overflow = _from_dense(nb_inner, neighbors.max_occupancy,
                                               nb_inner.cell_list_capacity)
    return _wrap(nb_inner, idx, max_occupancy, overflow, builds)

  return NeighborListFns(allocate_fn, update_fn)


_masked_neighbor_list.extra = 0


def neighbor_list_mask(neighbor: NeighborList, mask_self: bool = False):
  """partition.py:1167-1182."""
  N = len(neighbor.reference_position)
  if is_sparse(neighbor.format):
    mask = neighbor.idx[0] < N
    if mask_self:
      mask = mask & (neighbor.idx[0] != neighbor.idx[1])
    return mask
  mask = neighbor.idx < len(neighbor.idx)
  if mask_self:
    rows = torch.arange(N, dtype=torch.int32, device=neighbor.idx.device)
    mask = mask & (neighbor.idx != rows[:, None])
  return mask


class GraphsTuple(NamedTuple):
  """Field-for-field stand-in for `jraph.GraphsTuple` (jraph is not installed here): what
  `to_jraph` returns; pass its fields to `jraph.GraphsTuple(**g._asdict())` where jraph exists."""
  nodes: Any
  edges: Any
  receivers: Any
  senders: Any
  globals: Any
  n_node: Any
  n_edge: Any


def _tree_map(f, tree):
  if tree is None:
    return None
  if isinstance(tree, dict):
    return {k: _tree_map(f, v) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(_tree_map(f, v) for v in tree)
  return f(tree)


def to_jraph(neighbor: NeighborList, mask=None, nodes=None, edges=None, globals=None) -> GraphsTuple:
  """partition.py:1185-1243: sparse neighbour list -> graph tuple, padded by one fictitious
  graph with a single node (same edge reordering under an extra `mask`)."""
  if not is_sparse(neighbor.format):
    raise ValueError('Cannot convert a dense neighbor list to jraph format. Please use either '
                     'NeighborListFormat.Sparse or NeighborListFormat.OrderedSparse.')
  receivers, senders = neighbor.idx
  N = len(neighbor.reference_position)
  _mask = neighbor_list_mask(neighbor)

  def pad(x):
    return torch.cat((x, torch.zeros((1,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)), 0)
  nodes = _tree_map(pad, nodes)
  globals = _tree_map(pad, globals)
  if mask is not None:
    _mask = _mask & mask
    cumsum = torch.cumsum(_mask.to(torch.int64), 0)
    index = torch.where(_mask, cumsum - 1, torch.full_like(cumsum, len(receivers)))
    ordered = torch.full((len(receivers) + 1,), N, dtype=torch.int32, device=receivers.device)
    receivers = ordered.clone().index_put_((index,), receivers)[:-1]
    senders = ordered.clone().index_put_((index,), senders)[:-1]

    def reorder_edges(x):
      out = torch.zeros((len(x) + 1,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
      return out.index_put_((index,), x)[:-1]
    edges = _tree_map(reorder_edges, edges)
  n_edge = torch.stack([_mask.sum(), (~_mask).sum()])
  return GraphsTuple(nodes, edges, receivers, senders, globals,
                     torch.tensor([N, 1], device=receivers.device), n_edge)


def to_dense(neighbor: NeighborList):
  """partition.py:1245-1265 (host-side utility, torch ops)."""
  if neighbor.format is not Sparse:
    raise ValueError('Can only convert sparse neighbor lists to dense ones.')
  receivers, senders = neighbor.idx
  mask = neighbor_list_mask(neighbor)
  receivers, senders = receivers[mask].long(), senders[mask].long()
  N = len(neighbor.reference_position)
  count = torch.bincount(receivers, minlength=N)
  max_count = int(count.max())
  offset = torch.arange(max_count, device=receivers.device).repeat(N)[:len(senders)]
  hashes = senders * max_count + offset
  dense_idx = torch.full((N * max_count,), N, dtype=torch.int32,
                         device=receivers.device)
  dense_idx[hashes] = receivers.to(torch.int32)
  return dense_idx.reshape(N, max_count)
