import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jax_md_b200 as jmd
from tests import util
R_h, L = util.diamond(40, a=5.431, dtype=np.float32)
R_h = util.jitter(R_h, np.float32(L), 0.05, seed=1)
R = torch.as_tensor(R_h, device='cuda')
d, s = jmd.space.periodic(np.float32(L))
nf, efn = jmd.energy.stillinger_weber_neighbor_list(d, np.float32(L), capacity_multiplier=1.5)
nbrs = nf.allocate(R)
F = jmd.quantity.force(efn)
for _ in range(3): F(R, neighbor=nbrs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): F(R, neighbor=nbrs)
e1.record(); torch.cuda.synchronize()
print('SW force (pack + kernel) ms:', e0.elapsed_time(e1) / 20, 'max_occ', nbrs.max_occupancy)
