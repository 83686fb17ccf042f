"""ctypes wrapper of oracle/c/jmd_oracle.c: the multi-threaded CPU port used as
the CPU baseline and for larger-size parity.  TEST INFRASTRUCTURE ONLY.

Threading: the C file exposes [i0, i1) range functions; a Python thread pool
runs them concurrently (ctypes releases the GIL).  gcc in this image has no
libgomp, hence no OpenMP."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c')
_SO = os.path.join(_DIR, 'liboracle_c.so')
_lib = None


def build(force=False):
  src = os.path.join(_DIR, 'jmd_oracle.c')
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
    subprocess.run(['make', '-C', _DIR, '-B', 'liboracle_c.so'], check=True,
                   stdout=subprocess.DEVNULL)
  return _SO


def load():
  global _lib
  if _lib is None:
    build()
    lib = C.CDLL(_SO)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    lib.jo_nbr_create.restype = P
    lib.jo_nbr_create.argtypes = [I, F, F, F, I, I]
    lib.jo_nbr_free.argtypes = [P]
    lib.jo_nbr_bin.argtypes = [P, P]
    lib.jo_nbr_rows_range.argtypes = [P, P, I, I]
    lib.jo_nbr_rows_range.restype = I
    lib.jo_nbr_set_max_row.argtypes = [P, I]
    lib.jo_skin_range.argtypes = [P, P, I, I]
    lib.jo_skin_range.restype = I
    for f in ('jo_nbr_rows', 'jo_nbr_counts'):
      getattr(lib, f).restype = C.POINTER(C.c_int)
      getattr(lib, f).argtypes = [P]
    for f in ('jo_nbr_max_row', 'jo_nbr_max_cell', 'jo_nbr_overflow', 'jo_nbr_builds'):
      getattr(lib, f).restype = I
      getattr(lib, f).argtypes = [P]
    lib.jo_lj_force_range.restype = C.c_double
    lib.jo_lj_force_range.argtypes = [P, P, F, F, F, F, P, I, I]
    lib.jo_kick_drift_range.argtypes = [P, P, P, P, F, F, I, I]
    lib.jo_kick_range.argtypes = [P, P, F, I, I]
    _lib = lib
  return _lib


class LJSystem:
  """3-D periodic LJ (f32) with a full neighbour list, all on the CPU."""

  def __init__(self, R, L, r_cutoff=2.5, skin=0.3, r_onset=2.0, sigma=1.0, eps=1.0,
               row_capacity=128, dense=True, threads=None):
    self.lib = load()
    self.threads = threads or os.cpu_count() or 1
    self.pool = ThreadPoolExecutor(self.threads)
    self.R = np.ascontiguousarray(R, np.float32).copy()
    self.n = len(self.R)
    self.args = (np.float32(sigma), np.float32(eps), np.float32(r_onset), np.float32(r_cutoff))
    self.nb = self.lib.jo_nbr_create(self.n, np.float32(L), np.float32(r_cutoff), np.float32(skin),
                                     row_capacity, 1 if dense else 0)
    if not self.nb:
      raise ValueError('box too small for a cell list')
    self.m = row_capacity
    self.F = np.zeros_like(self.R)
    self.P = np.zeros_like(self.R)
    chunks = max(self.threads * 4, 1)
    edges = np.linspace(0, self.n, chunks + 1).astype(int)
    self.ranges = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    self.rebuild()

  def _par(self, fn):
    return list(self.pool.map(lambda r: fn(r[0], r[1]), self.ranges))

  def rebuild(self):
    Rp = self.R.ctypes.data
    self.lib.jo_nbr_bin(self.nb, Rp)
    mx = max(self._par(lambda a, b: self.lib.jo_nbr_rows_range(self.nb, Rp, a, b)))
    self.lib.jo_nbr_set_max_row(self.nb, mx)

  def update(self):
    Rp = self.R.ctypes.data
    if any(self._par(lambda a, b: self.lib.jo_skin_range(self.nb, Rp, a, b))):
      self.rebuild()
      return 1
    return 0

  def rows(self):
    ptr = self.lib.jo_nbr_rows(self.nb)
    return np.ctypeslib.as_array(ptr, shape=(self.n, self.m)).copy()

  def force(self):
    Rp, Fp = self.R.ctypes.data, self.F.ctypes.data
    e = sum(self._par(lambda a, b: self.lib.jo_lj_force_range(self.nb, Rp, *self.args, Fp, a, b)))
    return float(e), self.F.copy()

  def run(self, P, dt, steps, mass=1.0):
    """`steps` x (NeighborList.update + velocity Verlet); returns #rebuilds."""
    if P is not None:
      self.P = np.ascontiguousarray(P, np.float32).copy()
      self.force()
    Rp, Pp, Fp = self.R.ctypes.data, self.P.ctypes.data, self.F.ctypes.data
    dt, mass = np.float32(dt), np.float32(mass)
    rebuilds = 0
    for _ in range(steps):
      rebuilds += self.update()
      self._par(lambda a, b: self.lib.jo_kick_drift_range(self.nb, Rp, Pp, Fp, mass, dt, a, b))
      self._par(lambda a, b: self.lib.jo_lj_force_range(self.nb, Rp, *self.args, Fp, a, b))
      self._par(lambda a, b: self.lib.jo_kick_range(Pp, Fp, dt, a, b))
    return rebuilds

  @property
  def max_row(self):
    return self.lib.jo_nbr_max_row(self.nb)

  @property
  def overflow(self):
    return bool(self.lib.jo_nbr_overflow(self.nb))

  def close(self):
    self.pool.shutdown()
    if self.nb:
      self.lib.jo_nbr_free(self.nb)
      self.nb = None
