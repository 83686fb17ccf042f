"""Times the fused force + half-kick kernel (LJ fcc N = 4 n^3, f32, OrderedSparse list)
for the library selected by JMD_B200_LIB.  usage: python tools/probe_force.py [n] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, jax_md_b200 as jmd
from jax_md_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
R, box = bench.fcc((n, n, n)); L = box[0]; N = len(R)
rng = np.random.default_rng(5)
R = np.mod(R + rng.normal(0, 0.08, R.shape).astype(np.float32), L).astype(np.float32)
d, s = jmd.space.periodic(L)
nf, efn = jmd.energy.lennard_jones_neighbor_list(d, L, r_onset=2.0, r_cutoff=2.5, dr_threshold=0.3)
Rd = torch.as_tensor(R, device='cuda'); nb = nf.allocate(Rd)
init, step = jmd.simulate.nve(efn, s, 5e-3)
st = init(0, Rd, kT=1.0, momenta=torch.as_tensor(bench.momenta(N), device='cuda'), neighbor=nb)
for _ in range(30):
  nb = nb.update(st.position); st = step(st, neighbor=nb)
P = st.momentum.clone(); red = step._stepper.red(st.position)
_, sp, params = efn._resolve(nb, {})
ts = []
for i in range(reps + 5):
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  out = efn.launch(st.position, nb, sp, params, want_energy=False, momentum=P, mass=st.mass, dt_2=0.0, red=red,
                   refresh_positions=False)
  b.record(); torch.cuda.synchronize()
  if i >= 5: ts.append(a.elapsed_time(b))
F = out['force']
print('%-40s force+kick %.4f ms (min %.4f)  |F|sum %.6e' % (os.path.basename(_lib.LIB_PATH), np.mean(ts), np.min(ts),
                                                         float(F.double().abs().sum())))
