"""CPU-only checks of the XLA-FFI binding sources (csrc/jmd_ffi.cc, _jax_binding.py).  The
XLA headers and jax are absent from this image, so the handlers are (a) type-checked
against a stand-in of the FFI API (tests/ffi_stub) that static_asserts every handler's
C++ signature against its Bind() chain, (b) checked for coverage of every enqueue-only
entry point of include/jmd_b200.h, and (c) checked against the ctypes descriptor mirror
and the workspace table the torch host allocates from."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FFI = os.path.join(ROOT, 'jax_md_b200', 'csrc', 'jmd_ffi.cc')
HOST_ONLY = {'jmd_nbr_state_host', 'jmd_red_scratch_doubles', 'jmd_sw_scratch_ints', 'jmd_version', 'jmd_p2p_alloc',
             'jmd_p2p_open', 'jmd_p2p_close', 'jmd_p2p_free', 'jmd_host_flag_alloc', 'jmd_host_flag_free'}


def _declared():
  src = open(os.path.join(ROOT, 'include', 'jmd_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(jmd_[a-z0-9_]+)\s*\(', src)))


def test_every_enqueue_entry_point_has_a_handler():
  from jax_md_b200 import _jax_binding
  src = open(FFI).read()
  handlers = set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(jmd_ffi_[a-z0-9_]+)', src))
  want = {'jmd_ffi_' + n[len('jmd_'):] for n in _declared() if n not in HOST_ONLY}
  assert handlers == want, (sorted(want - handlers), sorted(handlers - want))
  assert {'jmd_ffi_' + h for h in _jax_binding.HANDLERS} == handlers
  # each handler forwards to the launcher of the same name
  for h in sorted(handlers):
    assert re.search(r'\b' + 'jmd_' + h[len('jmd_ffi_'):] + r'\(', src), h


def test_handlers_type_check_against_the_ffi_api():
  cmd = ['g++', '-std=c++17', '-fsyntax-only', '-I', os.path.join(ROOT, 'tests', 'ffi_stub'),
         '-I', os.path.join(ROOT, 'include'), '-I', '/usr/local/cuda/include', FFI]
  p = subprocess.run(cmd, capture_output=True, text=True)
  assert p.returncode == 0, p.stderr[-3000:]


def test_workspace_bundle_matches_descriptor_and_host_table():
  from jax_md_b200 import _jax_binding, _lib, partition
  src = open(FFI).read()
  macro = re.search(r'#define JMD_NBR_WORKSPACE\(X\)(.*?)\nconstexpr int kNbrWorkspace = (\d+);', src, re.S)
  fields = re.findall(r'X\((\w+)\)', macro.group(1))
  assert len(fields) == int(macro.group(2))
  assert tuple(fields) == _jax_binding.WORKSPACE
  pointer_fields = [n for n, t in _lib.NbrT._fields_ if t is C.c_void_p]
  # same relative order as the struct; the rest are optional (patched separately / NULL)
  assert [f for f in pointer_fields if f in fields] == fields
  assert set(pointer_fields) - set(fields) == {'species', 'nl16', 'blk_table', 'cs_bits', 'n_dev'}
  # every bundle member is a buffer the host allocates (nl / idx are sized after the occupancy pass)
  c = _lib.NbrT()
  c.n_pad, c.n_fine_cells = 32, 27
  table = {name for name, *_ in partition.workspace_buffers(c, 8, 3, 27)}
  assert set(fields) - table == {'nl', 'idx'}


def test_descriptor_serialisation_clears_pointers():
  from jax_md_b200 import _jax_binding, _lib
  nb = _lib.NbrT()
  nb.n, nb.m_int, nb.nl = 7, 3, 0xdeadbeef
  raw = _jax_binding._desc(nb)
  assert len(raw) == C.sizeof(_lib.NbrT)
  back = _lib.NbrT.from_buffer_copy(raw)
  assert back.n == 7 and back.m_int == 3 and not back.nl
