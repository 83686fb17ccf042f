"""Two gloo ranks on ONE GPU: trace the rebuild pipeline of the slab decomposition."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist, torch.multiprocessing as mp


def worker(rank, world, port):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  torch.cuda.set_device(0)
  import jax_md_b200 as jmd
  from jax_md_b200.domain import RingComm, SlabDomain
  from tests import test_domain as td
  R, P, box = td._system(np.float32)
  comm = RingComm()
  width = float(box[0]) / world
  own = np.floor(R[:, 0] / width).astype(int) % world == rank
  disp, shift = jmd.space.periodic(box)
  _, efn = jmd.energy.lennard_jones_neighbor_list(disp, box, dr_threshold=0.3)
  dom = SlabDomain(box, efn, 2.5, 0.3, 2e-3, comm=comm)
  gid = torch.as_tensor(np.nonzero(own)[0], device='cuda')
  st = dom.init(torch.as_tensor(R[own], device='cuda'), torch.as_tensor(P[own], device='cuda'), gid)
  print(rank, 'init', dom.last_info[:10], 'n_own', st.n_own, 'cap', dom.cap, dom.cap_list, dom.cap_mig, flush=True)
  last = dom.rebuilds
  for i in range(150):
    st = dom.step(st)
    if dom.rebuilds != last:
      last = dom.rebuilds
      x = st.position[:, 0]
      d = torch.remainder(x - dom.lo, float(box[0]))
      g = st.global_id
      print(rank, 'step', i, dom.last_info[:10], 'nan', bool(torch.isnan(st.R[:st.n_own + st.n_ghost]).any()),
            'outside', int((d >= dom.width).sum()), 'gid unique', g.unique().numel() == g.numel(),
            'gid min', int(g.min()), flush=True)
  tot = comm.sum(torch.tensor([float(st.n_own)]))
  print(rank, 'total atoms', float(tot), 'expected', len(R), 'ke', dom.kinetic_energy(), flush=True)
  dist.destroy_process_group()


if __name__ == '__main__':
  mp.spawn(worker, args=(2, 29533), nprocs=2, join=True)
