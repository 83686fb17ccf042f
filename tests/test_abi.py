"""CPU-only: the C-ABI library builds, loads and exports every symbol that
include/jmd_b200.h declares; host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, 'include', 'jmd_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(jmd_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  from jax_md_b200 import build, _lib
  path = build.build()
  lib = ctypes.CDLL(path)
  names = _declared()
  assert len(names) >= 14
  for n in names:
    assert hasattr(lib, n), f'{n} declared in jmd_b200.h but not exported'
  assert sorted(_lib.EXPORTED) == names, (sorted(_lib.EXPORTED), names)
  assert b'sm_100a' in _lib.load().jmd_version()


def test_struct_layout_matches_header():
  """ctypes mirrors of the PODs must have the C layout (sizes from the header
  arithmetic: no hidden padding surprises)."""
  from jax_md_b200 import _lib
  assert ctypes.sizeof(_lib.SpaceT) == 16 + 48 + 8 + 24 + 2 * 72
  assert ctypes.sizeof(_lib.SwT) == 64
  assert ctypes.sizeof(_lib.PairT) == 8 + 12 + 4 + 8 + 24 + 24 + 40
  n = _lib.NbrT
  assert n.n_pad.offset % 8 == 0 and n.space.offset % 8 == 0
  assert n.cell_count.offset == n.space.offset + ctypes.sizeof(_lib.SpaceT)


def test_no_cpu_fallback_without_cuda():
  import torch
  import jax_md_b200 as jmd
  if torch.cuda.is_available():
    pytest.skip('CUDA present')
  d, s = jmd.space.periodic(10.0)
  nf = jmd.partition.neighbor_list(d, 10.0, 1.0, 0.1)
  with pytest.raises(jmd._lib.JmdError):
    nf.allocate(torch.zeros(8, 3))


def test_host_capacity_rules_and_cell_dimensions():
  """partition.py:146-188 on the host."""
  from jax_md_b200 import partition
  box = np.float32(33.592)
  _, cs, cps, count = partition._cell_dimensions(3, box, np.float32(2.8))
  assert int(cps) == 11 and count == 1331
  assert abs(float(cs) - 3.0538) < 1e-3
  with pytest.raises(ValueError):
    partition._cell_dimensions(3, np.array([[5.0, 9.0, 9.0]], np.float32), np.float32(2.0))


def test_space_struct_values():
  import torch
  from jax_md_b200 import space
  d, s = space.periodic(np.float32(10.5))
  st = space.space_struct(d._jmd_space, 3, torch.float32)
  assert st.kind == 1 and st.wrapped == 1 and st.dim == 3
  assert st.side[0] == 10.5 and st.half[2] == 5.25
  d2, _ = space.free()
  assert space.space_struct(d2._jmd_space, 2, torch.float64).kind == 0
  # python-float side in x64: half is f32(side) * 0.5 (space.py:224 weak types)
  d3, _ = space.periodic(10.1)
  st = space.space_struct(d3._jmd_space, 3, torch.float64)
  assert st.side[0] == 10.1 and st.half[0] == float(np.float32(10.1) * np.float32(0.5))


def test_ctypes_layout_equals_the_c_compiler_layout(tmp_path):
  """Every field of the ctypes mirrors sits at the offset gcc gives the header's
  structs (guards the descriptor against silent drift when fields are added)."""
  import subprocess
  from jax_md_b200 import _lib
  structs = {'jmd_space_t': _lib.SpaceT, 'jmd_nbr_t': _lib.NbrT, 'jmd_pair_t': _lib.PairT,
             'jmd_sw_t': _lib.SwT}
  lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "jmd_b200.h"', 'int main(void) {']
  for cname, ct in structs.items():
    lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
    for fname, _ in ct._fields_:
      if fname.startswith('_pad'):
        continue
      lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
  lines += ['  return 0;', '}']
  src = tmp_path / 'layout.c'
  src.write_text('\n'.join(lines))
  exe = tmp_path / 'layout'
  subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
  out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
  got = dict(l.split() for l in out.splitlines())
  for cname, ct in structs.items():
    assert int(got[cname]) == ctypes.sizeof(ct), cname
    for fname, _ in ct._fields_:
      if fname.startswith('_pad'):
        continue
      assert int(got[f'{cname}.{fname}']) == getattr(ct, fname).offset, f'{cname}.{fname}'
