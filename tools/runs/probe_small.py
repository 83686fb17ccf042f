import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
R_h, box = bench.fcc((n, n, n)); N = len(R_h); L = box[0]
disp, shift = jmd.space.periodic(L)
nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, L, dr_threshold=bench.SKIN, capacity_multiplier=1.5)
init_fn, apply_fn = jmd.simulate.nve(efn, shift, bench.DT)
Rd = torch.as_tensor(R_h, device='cuda'); Pd = torch.as_tensor(bench.momenta(N), device='cuda')
nbrs = nf.allocate(Rd)
st = init_fn(0, Rd, kT=1.0, momenta=Pd, neighbor=nbrs)
for i in range(120):
  nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
torch.cuda.synchronize()
