"""Single-process run of the slab-domain step (world = 1: the peer-memory kernels talk to
this rank's own block) -- what ncu profiles for the k_dd_* kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jax_md_b200 import domain, space, energy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
R, box_l = bench.fcc((n, n, n))
box = box_l
disp, shift = space.periodic(box)
_, efn = energy.lennard_jones_neighbor_list(disp, box, r_onset=2.0, r_cutoff=bench.R_CUT, dr_threshold=bench.SKIN)
dom = domain.SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, use_graph=False)
st = dom.init(torch.as_tensor(R, device='cuda'), torch.as_tensor(bench.momenta(len(R)), device='cuda'))
for _ in range(30):
  st = dom.step(st)
torch.cuda.synchronize()
print('dd1 ok', st.n_own, dom.rebuilds, dom.kinetic_energy() / st.n_own)
dom.close()
