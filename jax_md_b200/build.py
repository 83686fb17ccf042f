"""Builds jax_md_b200/libjmd_b200.so in-tree with nvcc for sm_100a.

    python -m jax_md_b200.build [--force] [--verbose]

Every unit is an ordinary whole-program compile (no -rdc).  The persistent
neighbour-list update kernel synchronises its single co-resident wave with a hand-rolled
barrier (ordinary launch; see grid_sync in csrc/jmd_neighbor.cu for the time-out guard).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libjmd_b200.so')
OBJ = os.path.join(HERE, 'csrc', '_build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
          '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
          '--expt-relaxed-constexpr', '-Xptxas', '-v']
RDC_UNITS = []
UNITS = ['jmd_neighbor.cu', 'jmd_pair.cu', 'jmd_pair_staged.cu', 'jmd_pair_tric.cu', 'jmd_integrate.cu', 'jmd_sw.cu',
         'jmd_domain.cu']


def _sources():
  files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
           if f.endswith(('.cu', '.cuh', '.h'))]
  files.append(os.path.join(ROOT, 'include', 'jmd_b200.h'))
  files.append(os.path.abspath(__file__))
  return files


def _digest():
  h = hashlib.sha256()
  for f in _sources():
    with open(f, 'rb') as fh:
      h.update(f.encode())
      h.update(fh.read())
  return h.hexdigest()


def _run(cmd, verbose, log):
  p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                     text=True)
  log.write('$ ' + ' '.join(cmd) + '\n' + p.stdout + '\n')
  if verbose or p.returncode:
    sys.stderr.write('$ ' + ' '.join(cmd) + '\n' + p.stdout + '\n')
  if p.returncode:
    raise RuntimeError('nvcc failed: ' + ' '.join(cmd))


def build(force=False, verbose=False):
  """Compile every CUDA unit for sm_100a; returns the .so path."""
  os.makedirs(OBJ, exist_ok=True)
  stamp = OUT + '.digest'      # next to the .so: travels with it (csrc/_build does not)
  digest = _digest()
  if (not force and os.path.exists(OUT) and os.path.exists(stamp)
      and open(stamp).read() == digest):
    return OUT
  objs = []
  procs = []
  with open(os.path.join(OBJ, 'build.log'), 'w') as log:
    jobs = []
    for u in RDC_UNITS + UNITS:
      src = os.path.join(CSRC, u)
      if not os.path.exists(src):
        continue
      obj = os.path.join(OBJ, u.replace('.cu', '.o'))
      flag = ['-dc'] if u in RDC_UNITS else ['-c']
      jobs.append((u, [NVCC] + ARCH + COMMON + flag + [src, '-o', obj]))
      objs.append(obj)
    for u, cmd in jobs:      # compile units in parallel
      procs.append((u, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                             stderr=subprocess.STDOUT, text=True)))
    failed = None
    for u, cmd, p in procs:
      out, _ = p.communicate()
      log.write('$ ' + ' '.join(cmd) + '\n' + out + '\n')
      if verbose or p.returncode:
        sys.stderr.write('$ ' + ' '.join(cmd) + '\n' + out + '\n')
      if p.returncode:
        failed = u
    if failed:
      raise RuntimeError('nvcc failed on ' + failed)
    _run([NVCC, '-shared', '-o', OUT] + objs + ['-lcudart'], verbose, log)
  with open(stamp, 'w') as f:
    f.write(digest)
  return OUT


def ffi_include_dir():
  """Where the XLA FFI headers live ($JMD_XLA_FFI_INCLUDE or jax.ffi.include_dir()), or None."""
  d = os.environ.get('JMD_XLA_FFI_INCLUDE')
  if d and os.path.exists(os.path.join(d, 'xla', 'ffi', 'api', 'ffi.h')):
    return d
  try:
    import jax
    return jax.ffi.include_dir()
  except Exception:
    return None


def build_ffi(verbose=False):
  """libjmd_b200_ffi.so (csrc/jmd_ffi.cc) -- only where the XLA FFI headers exist; returns
  the path or None.  This image has no jax: the source is type-checked against a stand-in
  API by tests/test_ffi_binding.py instead."""
  inc = ffi_include_dir()
  if inc is None:
    return None
  out = os.path.join(HERE, 'libjmd_b200_ffi.so')
  cmd = ['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', inc, '-I', os.path.join(ROOT, 'include'),
         '-I', '/usr/local/cuda/include', os.path.join(CSRC, 'jmd_ffi.cc'), '-o', out,
         '-L', HERE, '-ljmd_b200', '-Wl,-rpath,$ORIGIN', '-L', '/usr/local/cuda/lib64', '-lcudart']
  with open(os.path.join(OBJ, 'build_ffi.log'), 'w') as log:
    _run(cmd, verbose, log)
  return out


if __name__ == '__main__':
  path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
  print(path)
