"""Time the phases of the distributed step / rebuild (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from jax_md_b200 import domain, space, energy, _lib

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK'])); dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
n = 63
R_loc, box_loc = bench.fcc((n, n, n)); a = box_loc[0] / n
R_loc[:, 0] += rank * n * a
box = np.array([world * n * a, n * a, n * a], np.float32)
P_loc = np.random.default_rng(1000 + rank).normal(0, 1.0, R_loc.shape).astype(np.float32)
comm = domain.RingComm()
disp, shift = space.periodic(box)
_, efn = energy.lennard_jones_neighbor_list(disp, box, dr_threshold=bench.SKIN)
dom = domain.SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, comm=comm)
st = dom.init(torch.as_tensor(R_loc, device=dev), torch.as_tensor(P_loc, device=dev))
for _ in range(200): st = dom.step(st)
torch.cuda.synchronize(); dist.barrier()

def sync():
  torch.cuda.synchronize(); dist.barrier()

acc = {'migrate': [], 'ghosts': [], 'nbr': [], 'halo': [], 'decision': [], 'force': [], 'step': []}
for rep in range(8):
  for _ in range(8): st = dom.step(st)
  dom._take_decision(st)
  sync(); t0 = time.perf_counter(); dom._migrate(st); sync(); t1 = time.perf_counter()
  dom._ghosts(st); sync(); t2 = time.perf_counter()
  ws = dom.nbrs._ws; ws.c.n, ws.c.n_rows = st.n_own + st.n_ghost, st.n_own; ws.n = ws.c.n
  s_, pp = _lib.stream(), _lib.ptr(st.R)
  _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, s_); _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 0, s_); _lib.call('jmd_nbr_export', ws.ref(), pp, 0, s_)
  sync(); t3 = time.perf_counter()
  dom._halo(st); sync(); t4 = time.perf_counter()
  dom._launch_decision(st); dom._take_decision(st); sync(); t5 = time.perf_counter()
  dom._force(st, kick=False); sync(); t6 = time.perf_counter()
  for k, v in zip(('migrate', 'ghosts', 'nbr', 'halo', 'decision', 'force'), (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5)):
    acc[k].append(v)
sync(); t0 = time.perf_counter()
for _ in range(100): st = dom.step(st)
sync(); acc['step'].append((time.perf_counter() - t0) / 100)
if rank == 0:
  print('us: ' + ' '.join(f'{k} {1e6*float(np.median(v)):.0f}' for k, v in acc.items()), 'rebuilds', dom.rebuilds)
dist.destroy_process_group()
