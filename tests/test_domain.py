"""Slab domain decomposition: ring plumbing on CPU (gloo,
world_size 2) and, on a GPU, two ranks sharing the device reproduce the
single-GPU trajectory."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _init(rank, world, port, backend='gloo'):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  if backend == 'nccl':
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
  else:
    dist.init_process_group('gloo', rank=rank, world_size=world)


# ---- CPU: ring exchange semantics ------------------------------------------------

def _ring_worker(rank, world, port, out):
  _init(rank, world, port)
  from jax_md_b200.domain import RingComm
  comm = RingComm()
  assert comm.left == (rank - 1) % world and comm.right == (rank + 1) % world
  # variable-size payloads: to the left 3+rank rows, to the right 5+rank rows
  n_l, n_r = 3 + rank, 5 + rank
  from_l, from_r = comm.exchange_counts(n_l, n_r)
  # what arrives from my left neighbour is what IT sent to ITS right
  assert from_l == 5 + comm.left and from_r == 3 + comm.right
  sl = torch.full((n_l, 2), 100.0 * rank + 1)          # tag: "sent left by rank"
  sr = torch.full((n_r, 2), 100.0 * rank + 2)          # tag: "sent right by rank"
  rl = torch.empty((from_l, 2))
  rr = torch.empty((from_r, 2))
  comm.exchange(sl, sr, rl, rr)
  assert torch.all(rl == 100.0 * comm.left + 2)
  assert torch.all(rr == 100.0 * comm.right + 1)
  flag = torch.tensor([1 if rank == world - 1 else 0])
  assert comm.any(flag) is True
  assert comm.any(torch.tensor([0])) is False
  tot = comm.sum(torch.tensor([float(rank + 1)]))
  assert float(tot) == world * (world + 1) / 2
  # empty messages in one direction must not deadlock
  el = torch.empty((0, 2))
  fl, fr = comm.exchange_counts(0, 2)
  comm.exchange(el, torch.ones((2, 2)), torch.empty((fl, 2)), torch.empty((fr, 2)))
  dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_ring_comm_gloo(world):
  mp.spawn(_ring_worker, args=(world, _free_port(), None), nprocs=world, join=True)


# ---- GPU: 2 ranks on one device == single-GPU run ---------------------------------

def _system(dtype, world=2):
  a = (4.0 / 0.8442) ** (1.0 / 3.0)
  cells = (5 * world, 5, 5)      # slabs of 5 cells along x: width 8.4 > 2 * 2.8
  basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
  g = np.stack(np.meshgrid(*[np.arange(c) for c in cells], indexing='ij'), -1).reshape(-1, 1, 3)
  R = ((g + basis[None]) * a).reshape(-1, 3)
  box = np.array([c * a for c in cells], np.float32)
  rng = np.random.default_rng(7)
  R = np.mod(R + rng.normal(0, 0.03, R.shape), box).astype(dtype)
  P = util.momenta(len(R), 3, kT=1.0, seed=3, dtype=dtype)
  return R, P, box


def _dd_worker(rank, world, port, steps, dtype_name, outdir, transport):
  # one GPU per rank (NCCL control plane) when the box has them, else all ranks share
  # device 0 (gloo control plane; the peer-memory exchange works through CUDA IPC either way)
  own_gpu = torch.cuda.device_count() >= world
  torch.cuda.set_device(rank if own_gpu else 0)
  _init(rank, world, port, 'nccl' if own_gpu else 'gloo')
  import jax_md_b200 as jmd
  from jax_md_b200.domain import RingComm, SlabDomain
  dtype = np.dtype(dtype_name).type
  R, P, box = _system(dtype, world)
  comm = RingComm()
  width = float(box[0]) / world
  own = np.floor(R[:, 0] / width).astype(int) % world == rank
  disp, shift = jmd.space.periodic(box)
  _, efn = jmd.energy.lennard_jones_neighbor_list(disp, box, dr_threshold=0.3)
  dom = SlabDomain(box, efn, 2.5, 0.3, 2e-3, comm=comm, transport=transport)
  gid = torch.as_tensor(np.nonzero(own)[0], device='cuda')
  st = dom.init(torch.as_tensor(R[own], device='cuda'), torch.as_tensor(P[own], device='cuda'), gid)
  pe0 = dom.potential_energy(st)
  for _ in range(steps):
    st = dom.step(st)
  ke = dom.kinetic_energy()
  pe = dom.potential_energy(st)
  np.savez(os.path.join(outdir, f'rank{rank}.npz'), gid=st.global_id.cpu().numpy(),
           R=st.position.cpu().numpy(), P=st.momentum.cpu().numpy(), ke=ke, pe=pe, pe0=pe0,
           rebuilds=dom.rebuilds, n_ghost=st.n_ghost, graph=int(dom._graph is not None))
  dom.close()
  dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('dtype_name,world,transport',
                         [('float64', 2, 'p2p'), ('float32', 2, 'p2p'), ('float64', 4, 'p2p'),
                          ('float64', 2, 'nccl')])
def test_slabs_match_single_gpu(tmp_path, dtype_name, world, transport):
  """2 / 4 ranks (their own GPUs when the box has them) reproduce the single-GPU
  trajectory: peer-memory halo + graph-captured step ('p2p') and the host-orchestrated
  torch.distributed fallback ('nccl')."""
  import jax_md_b200 as jmd
  dtype = np.dtype(dtype_name).type
  steps = 150
  R, P, box = _system(dtype, world)
  # single-GPU reference trajectory through the public API
  disp, shift = jmd.space.periodic(box)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, box, dr_threshold=0.3,
                                                   format=jmd.partition.Dense)
  Rd = torch.as_tensor(R, device='cuda')
  nbrs = nf.allocate(Rd)
  init_fn, apply_fn = jmd.simulate.nve(efn, shift, 2e-3)
  state = init_fn(0, Rd, kT=1.0, momenta=torch.as_tensor(P, device='cuda'), neighbor=nbrs)
  pe0_ref = float(efn(state.position, neighbor=nbrs))
  for _ in range(steps):
    nbrs = nbrs.update(state.position)
    state = apply_fn(state, neighbor=nbrs)
  assert not bool(nbrs.did_buffer_overflow)
  R_ref = state.position.cpu().numpy()
  P_ref = state.momentum.cpu().numpy()
  ke_ref = float(jmd.quantity.kinetic_energy(momentum=state.momentum, mass=state.mass))
  pe_ref = float(efn(state.position, neighbor=nbrs))

  mp.spawn(_dd_worker, args=(world, _free_port(), steps, dtype_name, str(tmp_path), transport),
           nprocs=world, join=True)
  parts = [np.load(tmp_path / f'rank{r}.npz') for r in range(world)]
  if transport == 'p2p':
    assert all(int(p['graph']) == 1 for p in parts), 'the step should have been graph-captured'
  gid = np.concatenate([p['gid'] for p in parts])
  assert sorted(gid.tolist()) == list(range(len(R)))          # every atom owned exactly once
  Rdd = np.zeros_like(R_ref)
  Pdd = np.zeros_like(P_ref)
  Rdd[gid] = np.concatenate([p['R'] for p in parts])
  Pdd[gid] = np.concatenate([p['P'] for p in parts])
  assert all(int(p['n_ghost']) > 0 for p in parts)
  assert int(parts[0]['rebuilds']) > 2
  tol = 1e-9 if dtype_name == 'float64' else 2e-3
  d = Rdd - R_ref
  d -= np.round(d / box) * box
  assert np.abs(d).max() < tol
  np.testing.assert_allclose(Pdd, P_ref, atol=tol * 10, rtol=0)
  rt = 1e-10 if dtype_name == 'float64' else 1e-5
  np.testing.assert_allclose(float(parts[0]['pe0']), pe0_ref, rtol=rt)
  np.testing.assert_allclose(float(parts[0]['ke']), ke_ref, rtol=max(rt, 1e-6 if dtype_name == 'float32' else rt) * 10)
  np.testing.assert_allclose(float(parts[0]['pe']), pe_ref, rtol=rt * 100)
