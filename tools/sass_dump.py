"""Writes profiles/<tag>_sass_<kernel>.txt for the f32 / 3-D instances of the kernels on the hot
path (cuobjdump -sass of the in-tree library; needs no GPU).

  python tools/sass_dump.py r02
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'jax_md_b200', 'libjmd_b200.so')
# (file tag, regex on the demangled function name) -- the instance the headline / config benches run
WANT = [
    ('k_pair_force_f32_3d_lj_scalar_kick', r'k_pair_force<float, 3, 0, true, 1, true>'),
    ('k_nbr_stencil_scan_f32_3d_ordered_filter', r'k_nbr_stencil_scan<float, 3, 2, true, 1, true>'),
    ('k_nbr_export_fin_f32_3d', r'k_nbr_export_fin<float, 3>'),
    ('k_nbr_offsets_f32_3d', r'k_nbr_offsets<float, 3>'),
    ('k_update_f32_3d', r'k_update<float, 3>'),
    ('k_kick_drift_f32_3d', r'k_kick_drift<float, 3>'),
    ('k_sw_f32', r'k_sw<float>'),
    ('k_sw_compact_f32', r'k_sw_compact<float>'),
    ('k_nhc_half_step_f32', r'k_nhc_half_step<float>'),
    ('k_fire_mix_f32', r'k_fire_mix<float>'),
    ('k_dd_comm_push_f32', r'k_dd_comm_push<float'),
    ('k_dd_comm_wait_f32', r'k_dd_comm_wait<float'),
    ('k_dd_select_ordered_f32', r'k_dd_select_ordered<float'),
]


def main(tag):
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  blocks = re.split(r'(?m)^\s*Function : ', sass)[1:]
  names = [b.split('\n', 1)[0].strip() for b in blocks]
  dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True, check=True).stdout.split('\n')
  for file_tag, rx in WANT:
    hit = [i for i, d in enumerate(dem[:len(names)]) if re.search(rx, d)]
    if not hit:
      print('missing', rx)
      continue
    i = hit[0]
    body = blocks[i]
    ops = re.findall(r'(?m)^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', body)
    hist = {}
    for o in ops:
      hist[o.split('.')[0]] = hist.get(o.split('.')[0], 0) + 1
    top = ', '.join(f'{k} {v}' for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:14])
    path = os.path.join(ROOT, 'profiles', f'{tag}_sass_{file_tag}.txt')
    with open(path, 'w') as f:
      f.write(f'// {dem[i]}\n// {len(ops)} SASS instructions; most frequent: {top}\n')
      f.write('Function : ' + body)
    print(os.path.basename(path), len(ops), 'instr')


if __name__ == '__main__':
  main(sys.argv[1] if len(sys.argv) > 1 else 'r02')
