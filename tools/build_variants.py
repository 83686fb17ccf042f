"""Builds experiment variants of libjmd_b200.so (extra -D flags on one unit),
as jax_md_b200/libjmd_b200_<tag>.so; select one with JMD_B200_LIB=<path>.

    python tools/build_variants.py tag unit.cu -DX=1 [-DY=2 ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_md_b200 import build as B   # noqa: E402


def main():
  tag, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
  B.build()
  obj = os.path.join(B.OBJ, unit.replace('.cu', f'_{tag}.o'))
  cmd = [B.NVCC] + B.ARCH + B.COMMON + flags + ['-c', os.path.join(B.CSRC, unit), '-o', obj]
  subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
  objs = [obj if u == unit else os.path.join(B.OBJ, u.replace('.cu', '.o')) for u in B.UNITS]
  out = os.path.join(B.HERE, f'libjmd_b200_{tag}.so')
  subprocess.run([B.NVCC, '-shared', '-o', out] + objs + ['-lcudart'], check=True)
  print(out)


if __name__ == '__main__':
  main()
