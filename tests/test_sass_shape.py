"""The gather loop of the headline force kernel is sensitive to how ptxas schedules it
("force kernel variants": the same source has come out 15-45 % slower after
unrelated edits).  The shape that measures 0.246 ms at N = 1M issues the four index loads of
an unrolled trip back to back as streaming loads (`LDG.E.EF`, from `__ldcs`) before the first
position gather (`LDG.E.128.CONSTANT`, from `__ldg`).  This test reads the SASS of the in-tree
library (cuobjdump; no GPU needed) and fails when a build loses that shape, so a slower kernel
cannot slip in unnoticed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'jax_md_b200', 'libjmd_b200.so')


@pytest.mark.skipif(shutil.which('cuobjdump') is None or not os.path.exists(LIB),
                    reason='needs cuobjdump and the built library')
def test_headline_force_kernel_issues_four_index_loads_before_the_first_gather():
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  blocks = re.split(r'(?m)^\s*Function : ', sass)[1:]
  names = [b.split('\n', 1)[0].strip() for b in blocks]
  dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True,
                       check=True).stdout.split('\n')
  hit = [i for i, d in enumerate(dem[:len(names)])
         if 'k_pair_force<float, 3, 0, true, 1, true>' in d]      # LJ, scalar params, reductions, kick
  assert len(hit) == 1, 'headline instance of k_pair_force not found'
  ops = re.findall(r'(?m)^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', blocks[hit[0]])
  loads = [o for o in ops if o.startswith('LDG')]
  first_gather = next(k for k, o in enumerate(loads) if o.startswith('LDG.E.128.CONSTANT'))
  streamed_before = [o for o in loads[:first_gather] if o.startswith('LDG.E.EF')]
  assert len(streamed_before) == 4, loads[:first_gather + 1]
  # no local-memory traffic: a spilled parameter struct was the 0.36 ms variant
  assert not any(o.startswith(('LDL', 'STL')) for o in ops)


@pytest.mark.skipif(shutil.which('cuobjdump') is None or not os.path.exists(LIB),
                    reason='needs cuobjdump and the built library')
def test_no_kernel_keeps_its_parameter_struct_in_local_memory():
  """A device function that takes the kernel-parameter struct by reference and is NOT inlined makes
  ptxas materialise the whole struct (0.9 - 1.2 KB) in local memory; every field read then goes
  through the stack (measured: the Dense stencil scan 0.8 -> 4.8 ms).  The stack size cuobjdump
  reports (kernel + callees) stays well under the struct size when that has not happened."""
  out = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True, check=True).stdout
  names = re.findall(r'Function ([^:\s]+):', out)
  stacks = [int(x) for x in re.findall(r'STACK:(\d+)', out)]
  assert names and len(names) == len(stacks)
  dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True,
                       check=True).stdout.split('\n')
  bad = [(d[:120], s) for d, s in zip(dem, stacks)
         if s >= (1000 if '<double' in d else 600)]
  assert not bad, bad
  # the pre-filtered f32 scans (every production rebuild: LJ / soft-sphere / SW lists, the domain
  # decomposition's Dense list) run without any stack at all; a CALL left in their loop body was
  # enough to cost the Dense one 6x
  hot = [(d[:120], s) for d, s in zip(dem, stacks)
         if re.search(r'k_nbr_stencil_scan<float, \d, \d, true, \d, true>', d) and s != 0]
  assert not hot, hot
  assert any(re.search(r'k_nbr_stencil_scan<float, 3, 0, true, 1, true>', d) for d in dem)
