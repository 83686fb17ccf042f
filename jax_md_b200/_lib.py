"""ctypes binding of libjmd_b200.so (include/jmd_b200.h).

PyTorch is used only as plumbing: device allocations, the current CUDA stream
and torch.distributed.  There is no CPU path: if the CUDA library is missing or
no CUDA device is visible, every compute entry point raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('JMD_B200_LIB', os.path.join(_HERE, 'libjmd_b200.so'))

F32, F64 = 0, 1
DENSE, SPARSE, ORDERED_SPARSE = 0, 1, 2
SPACE_FREE, SPACE_PERIODIC = 0, 1
POT_LJ, POT_SOFT_SPHERE, POT_MORSE = 0, 1, 2
PARAM_SCALAR, PARAM_PER_ATOM, PARAM_SPECIES, PARAM_MATRIX = 0, 1, 2, 3
ST_REBUILD, ST_MAX_CELL_OCC, ST_MAX_ROW, ST_TOTAL, ST_BUILDS = 0, 1, 2, 3, 4
ST_EXPORT_PENDING = 8
ST_COUNT = 16
# jmd_dd_* info words (include/jmd_b200.h)
(DD_N_OWN, DD_FACE_L, DD_FACE_R, DD_FROM_L, DD_FROM_R, DD_ERROR, DD_MIG_L, DD_MIG_R,
 DD_IN_L, DD_IN_R, DD_N_LOC, DD_N_ROWS) = range(12)
DD_INFO_COUNT = 16
DD_ELIST, DD_ECAP, DD_ETIMEOUT = 1, 2, 4
RED_ENERGY, RED_KINETIC, RED_VIRIAL = 0, 1, 2
RED_DSIGMA, RED_DEPSILON, RED_FF, RED_PP, RED_FP = 8, 9, 10, 11, 12
RED_COUNT = 16


class SpaceT(C.Structure):
  _fields_ = [('dim', C.c_int32), ('kind', C.c_int32), ('wrapped', C.c_int32),
              ('general', C.c_int32), ('side', C.c_double * 3),
              ('half', C.c_double * 3), ('fractional', C.c_int32), ('triclinic', C.c_int32),
              ('inv_box', C.c_double * 3), ('box_m', C.c_double * 9), ('inv_box_m', C.c_double * 9)]


class NbrT(C.Structure):
  _fields_ = [
      ('n', C.c_int32), ('dtype', C.c_int32), ('format', C.c_int32),
      ('use_cells', C.c_int32), ('mask_self', C.c_int32),
      ('always_rebuild', C.c_int32), ('cps', C.c_int32 * 3),
      ('n_cells', C.c_int32), ('cell_capacity', C.c_int32),
      ('m_int', C.c_int32), ('n_rows', C.c_int32), ('no_public_idx', C.c_int32),
      ('n_pad', C.c_int64), ('max_occupancy', C.c_int64),
      ('cell_size', C.c_double * 3), ('cutoff_sq', C.c_double),
      ('threshold_sq', C.c_double), ('space', SpaceT),
      ('cell_count', C.c_void_p), ('cell_start', C.c_void_p),
      ('cell_cursor', C.c_void_p), ('scan_tmp', C.c_void_p),
      ('hash', C.c_void_p), ('tmp_ids', C.c_void_p), ('perm', C.c_void_p),
      ('inv_perm', C.c_void_p), ('pos_sorted', C.c_void_p), ('nl', C.c_void_p),
      ('cnt', C.c_void_p), ('cnt_lower', C.c_void_p), ('offsets', C.c_void_p),
      ('reference_position', C.c_void_p), ('idx', C.c_void_p),
      ('error', C.c_void_p), ('state', C.c_void_p), ('species', C.c_void_p),
      ('fine_cps', C.c_int32 * 3), ('n_fine_cells', C.c_int32),
      ('stencil_w', C.c_int32), ('no_filter', C.c_int32),
      ('fine_cell_size', C.c_double * 3), ('ref_count', C.c_void_p),
      ('brick_shift', C.c_int32), ('staged', C.c_int32), ('ref_start', C.c_void_p),
      ('nl16', C.c_void_p), ('blk_table', C.c_void_p),
      ('skin_blk', C.c_void_p), ('skin_pre', C.c_int32), ('lazy_idx', C.c_int32),
      ('cell_scan', C.c_int32), ('cs_chunks', C.c_int32), ('cs_batches', C.c_int32),
      ('_pad2', C.c_int32), ('cs_bits', C.c_void_p), ('cs_lb', C.c_void_p),
      ('n_dev', C.c_void_p)]


DPARAM_MAX_SPECIES = 8        # JMD_DPARAM_MAX_SPECIES


class PairT(C.Structure):
  _fields_ = [('kind', C.c_int32), ('has_cutoff', C.c_int32),
              ('mode', C.c_int32 * 3), ('n_species', C.c_int32),
              ('transposed', C.c_int32), ('dparam_rows', C.c_int32),
              ('scalar', C.c_double * 3), ('array', C.c_void_p * 3),
              ('r_onset', C.c_double), ('r_cutoff', C.c_double),
              ('r_onset2', C.c_double), ('r_cutoff2', C.c_double),
              ('switch_denom', C.c_double)]


class DdT(C.Structure):
  """jmd_dd_t: per-step peer-memory exchange of the domain decomposition."""
  _fields_ = [('dtype', C.c_int32), ('dim', C.c_int32), ('rank', C.c_int32),
              ('world', C.c_int32), ('cap_list', C.c_int32),
              ('always_rebuild', C.c_int32),
              ('face_l', C.c_void_p), ('face_r', C.c_void_p),
              ('face_counts', C.c_void_p), ('info', C.c_void_p),
              ('epoch', C.c_void_p), ('ticket', C.c_void_p),
              ('skin_blk', C.c_void_p), ('land', C.c_void_p),
              ('signal', C.c_void_p), ('flags', C.c_void_p),
              ('peer_land_l', C.c_void_p), ('peer_land_r', C.c_void_p),
              ('peer_signal_l', C.c_void_p), ('peer_signal_r', C.c_void_p),
              ('peer_flags', C.c_void_p), ('host_flag', C.c_void_p)]


class SwT(C.Structure):
  _fields_ = [(k, C.c_double) for k in
              ('sigma', 'A', 'B', 'lam', 'gamma', 'epsilon',
               'three_body_strength', 'cutoff')]


_P, _I, _D, _L = C.c_void_p, C.c_int, C.c_double, C.c_int64
# name -> argtypes; every function returns int (0 == ok) unless noted.
_SIGNATURES = {
    'jmd_nbr_update': [C.POINTER(NbrT), _P, _P],
    'jmd_nbr_skin_check': [C.POINTER(NbrT), _P, _P],
    'jmd_nbr_bin': [C.POINTER(NbrT), _P, _I, _P],
    'jmd_nbr_build': [C.POINTER(NbrT), _P, _I, _I, _P],
    'jmd_nbr_export': [C.POINTER(NbrT), _P, _I, _P],
    'jmd_nbr_state_host': [C.POINTER(NbrT), C.POINTER(C.c_int64), _P],
    'jmd_nbr_pack': [C.POINTER(NbrT), _P, _P],
    'jmd_nbr_pack_range': [C.POINTER(NbrT), _P, _I, _I, _P],
    'jmd_dd_select': [_I, _I, _I, _P, _P, _I, _D, _D, _D, _D, _P, _P, _P, _I, _P],
    'jmd_dd_pack': [_I, _I, _I, _P, _P, _P, _P],
    'jmd_dd_pack_migrate': [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'jmd_dd_compact': [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'jmd_dd_pack_counted': [_I, _I, _I, _P, _P, _P, _P, _P],
    'jmd_dd_place': [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    'jmd_dd_select_ordered': [_I, _I, _I, _P, _P, _I, _D, _D, _D, _D, _P, _P, _P, _I, _P, _P],
    'jmd_p2p_alloc': [_L, C.POINTER(C.c_void_p), _P],
    'jmd_p2p_open': [_P, C.POINTER(C.c_void_p)],
    'jmd_p2p_close': [_P],
    'jmd_p2p_free': [_P],
    'jmd_host_flag_alloc': [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
    'jmd_host_flag_free': [_P],
    'jmd_dd_comm_push': [C.POINTER(DdT), _P, _P],
    'jmd_dd_comm_wait': [C.POINTER(DdT), C.POINTER(NbrT), _P, _P],
    'jmd_pair_force': [C.POINTER(NbrT), C.POINTER(PairT), _P, _P, _P, _P, _P,
                       _P, _P, _I, _D, _P, _I, _P],
    'jmd_sw_force': [C.POINTER(NbrT), C.POINTER(SwT), _P, _P, _P, _P, _P, _P, _I,
                     _D, _P, _P],
    'jmd_nve_kick_drift': [C.POINTER(SpaceT), _I, _I, C.POINTER(NbrT), _P, _P,
                           _P, _P, _I, _D, _P, _P, _P, _P, _P],
    'jmd_kick_reduce': [_I, _I, _I, _P, _P, _P, _I, _D, _P, _P, _P, _P],
    'jmd_scale_momentum': [_I, _L, _P, _P, _P],
    'jmd_nhc_half_step': [_I, _I, _I, _I, _D, _D, _L, _P, _P, _P, _P, _P, _P],
    'jmd_fire_mix': [_I, _L, _P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _D,
                     _D, _P],
}
EXPORTED = sorted(list(_SIGNATURES) + ['jmd_red_scratch_doubles', 'jmd_sw_scratch_ints', 'jmd_version'])

_lib = None


class JmdError(RuntimeError):
  pass


def load():
  """Loads the shared library; raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise JmdError(
        f'{LIB_PATH} is missing: build it with `python -m jax_md_b200.build` '
        '(there is no CPU fallback).')
  lib = C.CDLL(LIB_PATH)
  for name, args in _SIGNATURES.items():
    fn = getattr(lib, name)
    fn.argtypes = args
    fn.restype = C.c_int
  lib.jmd_red_scratch_doubles.argtypes = [C.c_int64]
  lib.jmd_red_scratch_doubles.restype = C.c_int64
  lib.jmd_sw_scratch_ints.argtypes = [C.POINTER(NbrT)]
  lib.jmd_sw_scratch_ints.restype = C.c_int64
  lib.jmd_version.restype = C.c_char_p
  _lib = lib
  return lib


def require_cuda():
  if not torch.cuda.is_available():
    raise JmdError('jax_md_b200 needs a CUDA device (sm_100a); none is visible '
                   'and there is no CPU fallback.')


def check(rc, what):
  if rc == 0:
    return
  if rc == -1:
    raise ValueError(f'{what}: invalid argument (JMD_EINVAL)')
  raise JmdError(f'{what}: CUDA error {rc}')


def stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
  if t is None:
    return None
  return C.c_void_p(t.data_ptr())


def dtype_code(dt):
  if dt == torch.float32:
    return F32
  if dt == torch.float64:
    return F64
  raise TypeError(f'positions must be float32 or float64, got {dt}')


def call(name, *args):
  fn = getattr(load(), name)
  check(fn(*args), name)
