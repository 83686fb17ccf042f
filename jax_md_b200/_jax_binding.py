"""JAX half of the XLA-FFI binding (csrc/jmd_ffi.cc): what a JAX MD maintainer adds to
use the B200 kernels from the reference's own plug-in points

    energy.lennard_jones_neighbor_list(displacement, box, ...,
        neighbor_list_fn=_jax_binding.neighbor_list,            # energy.py:312
        pair_neighbor_list_fn=_jax_binding.pair_neighbor_list)  # energy.py:313

`jax` is NOT installed in this image, so nothing here can run in this repo: the module is
import-safe without jax (everything touching jax is inside functions), the handler list
and the workspace bundle are checked against csrc/jmd_ffi.cc and include/jmd_b200.h by
tests/test_ffi_binding.py, and the torch-hosted mirror (`partition.py`, `smap.py`, ...)
is what the GPU tests exercise.  The host logic (capacity rules, descriptors) is shared
with that mirror: the descriptor structs are the same ctypes classes (`_lib.NbrT`, ...),
serialised with `bytes(struct)` into the handlers' byte-string attributes.
"""
import ctypes as C
import os

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB_PATH = os.environ.get('JMD_B200_FFI_LIB', os.path.join(_HERE, 'libjmd_b200_ffi.so'))

# One handler per C entry point that enqueues work (csrc/jmd_ffi.cc, symbol jmd_ffi_<name>).
HANDLERS = (
    'nbr_update', 'nbr_skin_check', 'nbr_bin', 'nbr_build', 'nbr_export', 'nbr_pack',
    'nbr_pack_range', 'pair_force', 'sw_force', 'nve_kick_drift', 'kick_reduce',
    'scale_momentum', 'nhc_half_step', 'fire_mix', 'dd_select', 'dd_select_ordered', 'dd_pack',
    'dd_pack_counted', 'dd_pack_migrate', 'dd_compact', 'dd_place', 'dd_comm_push',
    'dd_comm_wait')

# The neighbour-list workspace bundle, in the order jmd_ffi.cc patches the pointers
# (JMD_NBR_WORKSPACE).  (name, dtype, shape as a function of the static sizes)
WORKSPACE = (
    'cell_count', 'cell_start', 'cell_cursor', 'scan_tmp', 'hash', 'tmp_ids', 'perm',
    'inv_perm', 'pos_sorted', 'nl', 'cnt', 'cnt_lower', 'offsets', 'reference_position',
    'idx', 'error', 'state', 'ref_count', 'ref_start', 'skin_blk', 'cs_lb')


def register(lib_path=FFI_LIB_PATH):
  """jax.ffi.register_ffi_target for every handler (platform CUDA)."""
  import jax
  lib = C.CDLL(lib_path)
  for name in HANDLERS:
    jax.ffi.register_ffi_target('jmd_' + name, jax.ffi.pycapsule(getattr(lib, 'jmd_ffi_' + name)),
                                platform='CUDA')
  return lib


def _desc(struct):
  """Byte image of a POD descriptor with its pointer fields cleared (the handler patches
  them from XLA buffers)."""
  clone = type(struct).from_buffer_copy(struct)
  for name, ctype in clone._fields_:
    if ctype is C.c_void_p:
      setattr(clone, name, None)
  return bytes(clone)


def _ws_call(target, nb_struct, ws, operands, a0=0, a1=0):
  """ffi_call of a jmd_ffi_nbr_* handler: the workspace buffers are operands and results,
  aliased one to one so XLA updates them in place.  Returns the new workspace dict."""
  import jax
  bufs = [ws[k] for k in WORKSPACE]
  n_lead = len(operands)
  call = jax.ffi.ffi_call(
      target, [jax.ShapeDtypeStruct(b.shape, b.dtype) for b in bufs],
      input_output_aliases={n_lead + i: i for i in range(len(bufs))})
  out = call(*operands, *bufs, desc=_desc(nb_struct), a0=a0, a1=a1)
  return dict(zip(WORKSPACE, out))


def neighbor_list(displacement_or_metric, box, r_cutoff, dr_threshold=0.0,
                  capacity_multiplier=1.25, disable_cell_list=False, mask_self=True,
                  custom_mask_function=None, fractional_coordinates=False, format=None,
                  **static_kwargs):
  """Drop-in for `partition.neighbor_list` (partition.py:801-1164) returning the reference's
  own `NeighborListFns`.  `allocate` runs eagerly (it sizes buffers from occupancies read
  back to the host, exactly like the reference, partition.py:249,1094); `update` is ONE
  ffi_call of jmd_nbr_update -- the lax.cond of partition.py:1146 lives inside the kernel,
  every shape is static, overflow is a bit in `error.code`."""
  import jax
  import jax.numpy as jnp
  import numpy as np
  from jax_md import dataclasses as jdc
  from jax_md import partition as ref

  format = format or ref.NeighborListFormat.Dense
  if custom_mask_function is not None or fractional_coordinates:
    # not served by the kernels: keep the reference implementation for these lists
    return ref.neighbor_list(displacement_or_metric, box, r_cutoff, dr_threshold,
                             capacity_multiplier, disable_cell_list, mask_self, custom_mask_function,
                             fractional_coordinates, format, **static_kwargs)

  @jdc.dataclass
  class B200NeighborList(ref.NeighborList):
    """The reference pytree plus the hidden workspace (dynamic leaves) and the static
    descriptor the handlers need."""
    workspace: dict = None
    descriptor: bytes = jdc.static_field(default=b'')

  # host planning shared with the torch-hosted mirror: the same descriptor fill, buffer
  # table and capacity rule (jax_md_b200/partition.py)
  from . import partition as host
  from . import space as host_space
  spec = host_space.SpaceSpec(_lib.SPACE_PERIODIC, box, True)
  tagged = lambda Ra, Rb, **kw: displacement_or_metric(Ra, Rb, **kw)
  tagged._jmd_space = spec
  host_fns = host.neighbor_list(tagged, box, r_cutoff, dr_threshold, capacity_multiplier,
                                disable_cell_list, mask_self, format=host.NeighborListFormat[format.name],
                                **static_kwargs)
  fill = host_fns.allocate.fill_descriptor
  kinds = lambda dt: {'i4': jnp.int32, 'i8': jnp.int64, 'u1': jnp.uint8, 'f': dt}

  def allocate_fn(position, extra_capacity=0, **kwargs):
    N, dim = position.shape
    np_dtype = np.float32 if position.dtype == jnp.float32 else np.float64
    nb = _lib.NbrT()
    use_cells, cell_size, n_cells, _ = fill(nb, N, dim, np_dtype, N)
    nb.cell_capacity, nb.m_int, nb.max_occupancy = 1, 1, 1
    ws = {}
    for name, shape, kind, fill_value in host.workspace_buffers(nb, N, dim, n_cells):
      ws[name] = jnp.full(shape, 0 if fill_value is None else fill_value, kinds(position.dtype)[kind])
    ws['nl'] = jnp.zeros((1, nb.n_pad), jnp.int32)
    ws['idx'] = jnp.zeros((0,), jnp.int32)
    ws = _ws_call('jmd_nbr_bin', nb, ws, [position], a0=0)
    cl_capacity, width = None, N
    if use_cells:
      max_cell = int(ws['state'][_lib.ST_MAX_CELL_OCC])           # host sync, like the reference (:249)
      cl_capacity = int(max_cell * capacity_multiplier) + extra_capacity
      nb.cell_capacity = cl_capacity
      width = 3 ** dim * cl_capacity
      ws = _ws_call('jmd_nbr_bin', nb, ws, [position], a0=0)       # slot order depends on the capacity
    ws = _ws_call('jmd_nbr_build', nb, ws, [position], a0=1, a1=0)  # occupancy pass (:1083-1088)
    max_row, total = int(ws['state'][_lib.ST_MAX_ROW]), int(ws['state'][_lib.ST_TOTAL])
    hfmt = host.NeighborListFormat[format.name]
    max_occupancy, m_int = host.capacity_rule(hfmt, N, width, mask_self, capacity_multiplier,
                                              extra_capacity, max_row, total)
    nb.max_occupancy, nb.m_int = max_occupancy, m_int
    ws['nl'] = jnp.zeros((m_int, nb.n_pad), jnp.int32)
    idx_shape = (2, max_occupancy) if host.is_sparse(hfmt) else (N, max_occupancy)
    ws['idx'] = jnp.full(idx_shape, N, jnp.int32)
    ws = _ws_call('jmd_nbr_build', nb, ws, [position], a0=0, a1=0)
    ws = _ws_call('jmd_nbr_export', nb, ws, [position], a0=0)
    return B200NeighborList(ws['idx'], ws['reference_position'], ref.PartitionError(ws['error']),
                            cl_capacity, max_occupancy, format, cell_size, None,
                            update_fn, workspace=ws, descriptor=_desc(nb))

  def update_fn(position, neighbors, **kwargs):
    nb = _lib.NbrT.from_buffer_copy(neighbors.descriptor)
    ws = _ws_call('jmd_nbr_update', nb, neighbors.workspace, [position])
    return jdc.replace(neighbors, idx=ws['idx'], reference_position=ws['reference_position'],
                       error=ref.PartitionError(ws['error']), workspace=ws)

  return ref.NeighborListFns(allocate_fn, update_fn)


def pair_neighbor_list(fn, displacement_or_metric, species=None, reduce_axis=None,
                       ignore_unused_parameters=False, **kwargs):
  """Drop-in for `smap.pair_neighbor_list` (smap.py:856-979) for the potentials the fused
  kernel knows (tagged `_jmd_potential`); anything else falls back to the reference's
  generic implementation.  jax.grad w.r.t. positions, sigma and epsilon is a custom_vjp
  whose forward is ONE ffi_call of jmd_pair_force (E, F, dE/dsigma, dE/depsilon, virial)."""
  import jax
  import jax.numpy as jnp
  from jax_md import smap as ref_smap

  pot = getattr(fn, '_jmd_potential', None)
  if pot is None or reduce_axis is not None or species is not None:
    return ref_smap.pair_neighbor_list(fn, displacement_or_metric, species=species,
                                       reduce_axis=reduce_axis,
                                       ignore_unused_parameters=ignore_unused_parameters, **kwargs)
  kwargs.pop('fractional_coordinates', None)                     # energy.py:240,340,443

  def _launch(R, neighbor, sigma, epsilon):
    from . import smap as host
    nb = _lib.NbrT.from_buffer_copy(neighbor.descriptor)
    pt = host.pair_descriptor(pot, float(sigma), float(epsilon), kwargs.get('alpha'))     # _lib.PairT
    ws = neighbor.workspace
    bufs = [ws[k] for k in WORKSPACE]
    N, dim = R.shape
    outs = [jax.ShapeDtypeStruct((N, dim), R.dtype), jax.ShapeDtypeStruct((0,), R.dtype),
            jax.ShapeDtypeStruct((_lib.RED_COUNT,), jnp.float64), jax.ShapeDtypeStruct((0,), jnp.float64),
            jax.ShapeDtypeStruct((2 * (N // 128 + 2) * 16 + 8,), jnp.float64),
            jax.ShapeDtypeStruct((0,), R.dtype)]
    outs += [jax.ShapeDtypeStruct(b.shape, b.dtype) for b in bufs]
    empty = jnp.zeros((0,), R.dtype)
    lead = [empty, empty, empty, empty, empty, jnp.zeros((0,), jnp.int32), empty]
    call = jax.ffi.ffi_call('jmd_pair_force', outs,
                            input_output_aliases={len(lead) + i: 6 + i for i in range(len(bufs))})
    res = call(*lead, *bufs, nbr=_desc(nb), pair=_desc(pt), mass_is_array=0, dt_2=0.0,
               want_energy=1, kick=0)
    force, red = res[0], res[2]
    return red[_lib.RED_ENERGY].astype(R.dtype), force, red

  @jax.custom_vjp
  def energy(R, neighbor, sigma, epsilon):
    return _launch(R, neighbor, sigma, epsilon)[0]

  def fwd(R, neighbor, sigma, epsilon):
    E, F, red = _launch(R, neighbor, sigma, epsilon)
    return E, (F, red[_lib.RED_DSIGMA], red[_lib.RED_DEPSILON])

  def bwd(res, ct):
    F, dsig, deps = res
    return (-ct * F, None, (ct * dsig).astype(F.dtype), (ct * deps).astype(F.dtype))

  energy.defvjp(fwd, bwd)

  def energy_fn(R, neighbor=None, **dynamic_kwargs):
    sigma = dynamic_kwargs.get('sigma', kwargs.get('sigma', 1.0))
    epsilon = dynamic_kwargs.get('epsilon', kwargs.get('epsilon', 1.0))
    return energy(R, neighbor, sigma, epsilon)

  return energy_fn
