"""The C CPU port (oracle/c) against the NumPy oracle (CPU), and -- on a GPU --
bit-exact neighbour sets of the B200 path against the C port at N = 256,000."""
import numpy as np
import pytest

from oracle import cport
from oracle import energy as oenergy
from oracle import partition as opart
from oracle import simulate as osim
from oracle import space as ospace
from tests import util


def test_c_port_matches_numpy_oracle():
  R, L = util.fcc(9)
  R = util.jitter(R, L, 0.06)
  d, s = ospace.periodic(L)
  for dense in (True, False):
    sysc = cport.LJSystem(R, L, dense=dense)
    fmt = opart.Dense if dense else opart.Sparse
    nf = opart.neighbor_list(d, L, np.float32(2.5), np.float32(0.3), format=fmt)
    nb = nf.allocate(R)
    rows = sysc.rows()
    if dense:
      assert sysc.max_row == nb.occupancy
      np.testing.assert_array_equal(np.sort(rows[:, :nb.idx.shape[1]], -1), np.sort(nb.idx, -1))
    else:
      N = len(R)
      m = rows < N
      senders = np.broadcast_to(np.arange(N)[:, None], rows.shape)[m]
      pairs = np.stack([senders, rows[m]], 1)
      pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
      np.testing.assert_array_equal(pairs, util.sparse_pairs(nb.idx, N))
    sysc.close()
  sysc = cport.LJSystem(R, L)
  nf = opart.neighbor_list(d, L, np.float32(2.5), np.float32(0.3), format=opart.Dense)
  nb = nf.allocate(R)
  pot = oenergy.PairPotential('lj', np.float32(2.0), np.float32(2.5))
  E_o, F_o, _ = oenergy.pair_neighbor_list_energy(
      pot, d, R.astype(np.float64), nb, want_grads=True, sigma=np.float64(1), epsilon=np.float64(1))
  e, F = sysc.force()
  np.testing.assert_allclose(e, E_o, rtol=2e-5)
  np.testing.assert_allclose(F, F_o, rtol=1e-4, atol=1e-4 * np.abs(F_o).max())
  # trajectories: 40 steps of update + velocity Verlet
  P = util.momenta(len(R), 3, 1.0)
  holder = {'nb': nb}

  def force(Rx):
    holder['nb'] = holder['nb'].update(Rx)
    return oenergy.pair_neighbor_list_energy(pot, d, Rx, holder['nb'], want_grads=True,
                                             sigma=np.float32(1), epsilon=np.float32(1))[1]
  init, step = osim.nve(force, s, 5e-3)
  st = init(R, P, mass=np.float32(1.0))
  for _ in range(40):
    st = step(st)
  sysc.run(P, 5e-3, 40)
  dR = sysc.R - st.position
  dR -= np.round(dR / L) * L
  assert np.abs(dR).max() < 2e-4
  sysc.close()


@pytest.mark.gpu
@pytest.mark.parametrize('dense', [True, False])
def test_gpu_neighbor_sets_match_c_port_256k(dense):
  import torch
  import jax_md_b200 as jmd
  import bench
  R, box = bench.fcc((40, 40, 40))
  rng = np.random.default_rng(9)
  L = box[0]
  R = np.mod(R + rng.normal(0, 0.08, R.shape).astype(np.float32), L).astype(np.float32)
  N = len(R)
  sysc = cport.LJSystem(R, L, dense=dense)
  rows = sysc.rows()
  d, _ = jmd.space.periodic(L)
  fmt = jmd.partition.Dense if dense else jmd.partition.Sparse
  nf = jmd.partition.neighbor_list(d, L, np.float32(2.5), np.float32(0.3), format=fmt)
  nbrs = nf.allocate(torch.as_tensor(R, device='cuda'))
  assert int(nbrs.error.code) == 0
  idx = nbrs.idx.cpu().numpy()
  if dense:
    assert nbrs.max_occupancy == int(sysc.max_row * 1.25)
    m = idx.shape[1]
    assert sysc.max_row <= m
    np.testing.assert_array_equal(np.sort(idx, -1), np.sort(rows[:, :m], -1))
  else:
    mk = rows < N
    senders = np.broadcast_to(np.arange(N)[:, None], rows.shape)[mk]
    pairs = np.stack([senders, rows[mk]], 1).astype(np.int64)
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    np.testing.assert_array_equal(util.sparse_pairs(idx, N), pairs)
  sysc.close()
