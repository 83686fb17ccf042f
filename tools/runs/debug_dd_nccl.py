"""torchrun, NCCL: trace the rebuild pipeline of the slab decomposition on the bench system."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from jax_md_b200 import domain, space, energy

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK'])); dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
R_loc, box_loc = bench.fcc((n, n, n)); a = box_loc[0] / n
R_loc[:, 0] += rank * n * a
box = np.array([world * n * a, n * a, n * a], np.float32)
P_loc = np.random.default_rng(1000 + rank).normal(0, np.sqrt(bench.KT), R_loc.shape).astype(np.float32)
comm = domain.RingComm()
disp, shift = space.periodic(box)
_, efn = energy.lennard_jones_neighbor_list(disp, box, dr_threshold=bench.SKIN)
dom = domain.SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, comm=comm)
gid = torch.arange(len(R_loc), device=dev) + rank * len(R_loc)
st = dom.init(torch.as_tensor(R_loc, device=dev), torch.as_tensor(P_loc, device=dev), gid)


def report(tag):
  g = st.global_id
  x = st.position[:, 0]
  d = torch.remainder(x - dom.lo, float(box[0]))
  occ = dom.nbrs._ws.state_host()
  print(rank, tag, dom.last_info[:10], 'nan', bool(torch.isnan(st.R[:st.n_own + st.n_ghost]).any()),
        'outside', int((d >= dom.width).sum()), 'gid unique', g.unique().numel() == g.numel(),
        'gid range', int(g.min()), int(g.max()), 'state', list(occ)[:5],
        'Fmax', float(st.force.abs().max()), 'Pmax', float(st.momentum.abs().max()), flush=True)


report('init')
last = dom.rebuilds
for i in range(steps):
  st = dom.step(st)
  if dom.rebuilds != last:
    last = dom.rebuilds
    occ = dom.nbrs._ws.state_host()
    if rank == 0 and (last % 10 == 0 or occ[1] > dom.nbrs._ws.c.cell_capacity or occ[2] > dom.nbrs._ws.c.m_int):
      print('caps', dom.nbrs._ws.c.cell_capacity, dom.nbrs._ws.c.m_int, flush=True)
      report(f'step {i}')
tot = comm.sum(torch.tensor([float(st.n_own)], device=dev))
print(rank, 'total atoms', float(tot), 'expected', world * len(R_loc), 'ke/N', dom.kinetic_energy() / float(tot), flush=True)
dist.destroy_process_group()
