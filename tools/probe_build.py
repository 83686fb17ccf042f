"""Times bin / build / export of a forced rebuild for a few descriptor variants (one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
from jax_md_b200 import _lib


def timed(ws, pp, tag):
  st = _lib.stream()
  out = []
  for name, args in (('jmd_nbr_bin', (0,)), ('jmd_nbr_build', (0, 0)), ('jmd_nbr_export', (0,))):
    ts = []
    for _ in range(3):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      _lib.call(name, ws.ref(), pp, *args, st)
      b.record()
      torch.cuda.synchronize()
      ts.append(a.elapsed_time(b))
    out.append('%s %.3f' % (name[8:], min(ts)))
  s = ws.state_host()
  print(tag, ' | '.join(out), '| filter-relevant: cps', list(ws.c.cps), 'n', ws.c.n, 'n_rows', ws.c.n_rows,
        'm_int', ws.c.m_int, 'max_row', s[_lib.ST_MAX_ROW], 'err', int(ws.t['error'].item()), flush=True)


def run(cells, tag, jitter=0.0, fmt=None, **alloc_kw):
  R_h, box = bench.fcc(cells)
  if jitter:
    R_h = np.mod(R_h + np.random.default_rng(0).normal(0, jitter, R_h.shape).astype(np.float32), box).astype(np.float32)
  b = box[0] if cells[0] == cells[1] == cells[2] else box
  disp, shift = jmd.space.periodic(b)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, b, r_onset=2.0, r_cutoff=bench.R_CUT, dr_threshold=bench.SKIN,
                                                   **({} if fmt is None else {'format': fmt}))
  R = torch.as_tensor(R_h, device='cuda')
  nb = nf.allocate(R, **alloc_kw)
  timed(nb._ws, _lib.ptr(R), tag)
  return nb, R


run((63, 63, 63), 'cubic 1M lattice       ')
run((63, 63, 63), 'cubic 1M Dense         ', fmt=jmd.partition.Dense)
run((63, 63, 63), 'cubic 1M Sparse        ', fmt=jmd.partition.Sparse)
run((63, 63, 63), 'cubic 1M jitter 0.1    ', jitter=0.1)
run((126, 63, 63), 'slab box 2M lattice    ')
nb, R = run((63, 63, 63), 'cubic 1M n_rows=0.95n  ', n_rows=950000, n_capacity=1200000, no_public_idx=True)
# positions slightly outside the box (what a slab's atoms look like between wraps?)
R2 = R.clone(); R2[::50, 0] -= 1e-3
ws = nb._ws
timed(ws, _lib.ptr(R2), 'same, 2% of atoms at x<0')
