"""Per-step device time of the domain-decomposed LJ step (torchrun, >= 2 ranks):
  torchrun --nproc-per-node 2 tools/probe_dd_steps.py [cells] [steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jax_md_b200 import energy, space  # noqa: E402
from jax_md_b200.domain import RingComm, SlabDomain  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
R_loc, box_loc = bench.fcc((n, n, n))
a = box_loc[0] / n
R_loc[:, 0] += rank * n * a
box = np.array([world * n * a, n * a, n * a], np.float32)
N_loc = len(R_loc)
rng = np.random.default_rng(1000 + rank)
P_loc = rng.normal(0, np.sqrt(bench.KT), (N_loc, 3)).astype(np.float32)
comm = RingComm()
disp, shift = space.periodic(box)
_, efn = energy.lennard_jones_neighbor_list(disp, box, r_onset=2.0, r_cutoff=bench.R_CUT, dr_threshold=bench.SKIN)
dom = SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, comm=comm)
Rd, Pd = torch.as_tensor(R_loc, device=dev), torch.as_tensor(P_loc, device=dev)
Pd -= (comm.sum(Pd.sum(0, dtype=torch.float64)) / (world * N_loc)).to(Pd.dtype)
st = dom.init(Rd, Pd, torch.arange(N_loc, device=dev) + rank * N_loc)
# phase timing of the host-driven rebuild: CUDA events around every library call / exchange
from jax_md_b200 import _lib as L
phases = []
_call = L.call


def timed_call(name, *a):
  if torch.cuda.is_current_stream_capturing() or not (name.startswith('jmd_nbr_') or name.startswith('jmd_dd_')):
    return _call(name, *a)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  r = _call(name, *a)
  e1.record()
  phases.append((name, e0, e1))
  return r


L.call = timed_call
for meth in ('exchange', 'exchange_many'):
  def wrap(f, meth=meth):
    def g(*a, **k):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      r = f(*a, **k)
      e1.record()
      phases.append(('comm.' + meth, e0, e1))
      return r
    return g
  setattr(comm, meth, wrap(getattr(comm, meth)))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
reb, nown, ngh = [], [], []
import time
host = []
evs[0].record()
for i in range(steps):
  t0 = time.perf_counter()
  st = dom.step(st)
  host.append((time.perf_counter() - t0) * 1e3)
  evs[i + 1].record()
  reb.append(dom.rebuilds)
  nown.append(st.n_own)
  ngh.append(st.n_ghost)
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
if rank == 0:
  for i in range(steps):
    mark = 'R' if (i > 0 and reb[i] != reb[i - 1]) else ' '
    print('step %3d %s dev %.3f ms host %.3f ms own %d ghost %d' % (i, mark, ms[i], host[i], nown[i], ngh[i]))
  print('overflow', int(dom.nbrs.error.code), 'graph', dom._graph is not None)
  ws = dom.nbrs._ws
  n_loc = st.n_own + st.n_ghost
  Rl = st.R[:n_loc]
  cur = ws.t['cell_cursor']
  print('dirty cells', int(((cur & (1 << 30)) != 0).sum()), 'of', ws.c.n_cells, 'cps', list(ws.c.cps), 'cell_size',
        list(ws.c.cell_size), 'n', ws.c.n, 'n_rows', ws.c.n_rows, 'm_int', ws.c.m_int, 'no_filter', ws.c.no_filter,
        'use_cells', ws.c.use_cells, 'general', ws.c.space.general, 'tric', ws.c.space.triclinic)
  print('R min', Rl.min(0).values.tolist(), 'max', Rl.max(0).values.tolist(), 'box', box.tolist(),
        'side', list(ws.c.space.side), 'negative coords', int((Rl < 0).sum()), 'beyond box',
        int((Rl >= torch.as_tensor(box, device=dev)).sum()))
  agg = {}
  for name, e0, e1 in phases:
    agg.setdefault(name, []).append(e0.elapsed_time(e1))
  for name, v in agg.items():
    print('phase %-28s n %3d  mean %.3f ms  max %.3f ms' % (name, len(v), sum(v) / len(v), max(v)))
dom.close()
dist.destroy_process_group()
