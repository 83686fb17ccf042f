"""Measured facts for DESIGN.md 6: rebuild time of periodic_general lists (orthorhombic with the
pre-filter vs triclinic on the exact metric) and the cost of a host-orchestrated NPT step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
from jax_md_b200 import _lib


def build_ms(ws, pp):
  st = _lib.stream()
  ts = []
  for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, st)
    _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 0, st)
    _lib.call('jmd_nbr_export', ws.ref(), pp, 0, st)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return min(ts)


n = 40
R_h, box = bench.fcc((n, n, n))
L = float(box[0])
rng = np.random.default_rng(0)
S = np.mod((R_h + rng.normal(0, 0.03, R_h.shape)) / L, 1.0).astype(np.float32)
Sd = torch.as_tensor(S, device='cuda')
for tag, H in (('orthorhombic (vector box)', np.full(3, L, np.float32)),
               ('triclinic (matrix box)  ', np.array([[L, 0.2 * L, 0.1 * L], [0, L, 0.15 * L], [0, 0, L]], np.float32))):
  d, s = jmd.space.periodic_general(H)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(d, H, r_onset=2.0, r_cutoff=2.5, dr_threshold=0.3,
                                                   fractional_coordinates=True)
  nb = nf.allocate(Sd)
  print(tag, 'N', len(S), 'rebuild %.3f ms' % build_ms(nb._ws, _lib.ptr(Sd)), 'cells', list(nb._ws.c.cps),
        'overflow', bool(nb.did_buffer_overflow), flush=True)
  init, step = jmd.simulate.nve(efn, s, 5e-3)
  st = init(0, Sd, kT=1.0, neighbor=nb)
  for _ in range(20):
    nb = nb.update(st.position); st = step(st, neighbor=nb)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(100):
    nb = nb.update(st.position); st = step(st, neighbor=nb)
  torch.cuda.synchronize()
  print(tag, 'NVE eager %.3f ms/step' % ((time.perf_counter() - t0) * 10), flush=True)
  del nb, nf

# NPT, LJ, N = 256k, unit-cube coordinates
d, s = jmd.space.periodic_general(np.float32(L))
nf, efn = jmd.energy.lennard_jones_neighbor_list(d, np.float32(L), r_onset=2.0, r_cutoff=2.5, dr_threshold=0.3,
                                                 fractional_coordinates=True, format=jmd.partition.Dense,
                                                 capacity_multiplier=1.5)
nb = nf.allocate(Sd)
init, step = jmd.simulate.npt_nose_hoover(efn, s, 2e-3, 1.0, 1.0)
st = init(0, Sd, np.float32(L), neighbor=nb)
for _ in range(10):
  nb = nb.update(st.position, box=jmd.simulate.npt_box(st)); st = step(st, neighbor=nb)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100):
  nb = nb.update(st.position, box=jmd.simulate.npt_box(st)); st = step(st, neighbor=nb)
torch.cuda.synchronize()
print('NPT Nose-Hoover LJ N=%d: %.3f ms/step (host-orchestrated, one box read per step), box/L0 = %.4f' % (
    len(S), (time.perf_counter() - t0) * 10, float(jmd.simulate.npt_box(st)[0, 0]) / L), flush=True)
