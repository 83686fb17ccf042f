"""Builds experiment variants of libjmd_b200.so (extra -D flags on the pair-force
units), as jax_md_b200/libjmd_b200_<tag>.so; select one with JMD_B200_LIB=<path>.

    python tools/build_variants.py tag [unit.cu ...] -DX=1 [-DY=2 ...]

Units default to the pair-force units.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_md_b200 import build as B   # noqa: E402

PAIR_UNITS = ['jmd_pair.cu', 'jmd_pair_staged.cu']


def main():
  tag = sys.argv[1]
  units = [a for a in sys.argv[2:] if a.endswith('.cu')] or PAIR_UNITS
  flags = [a for a in sys.argv[2:] if not a.endswith('.cu')]
  B.build()
  procs, objs = [], []
  for u in B.UNITS:
    if u in units:
      obj = os.path.join(B.OBJ, u.replace('.cu', f'_{tag}.o'))
      cmd = [B.NVCC] + B.ARCH + B.COMMON + flags + ['-c', os.path.join(B.CSRC, u), '-o', obj]
      procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    else:
      obj = os.path.join(B.OBJ, u.replace('.cu', '.o'))
    objs.append(obj)
  for p in procs:
    assert p.wait() == 0
  out = os.path.join(B.HERE, f'libjmd_b200_{tag}.so')
  subprocess.run([B.NVCC, '-shared', '-o', out] + objs + ['-lcudart'], check=True,
                 stderr=subprocess.DEVNULL)
  print(out)


if __name__ == '__main__':
  main()
