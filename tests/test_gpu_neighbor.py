"""GPU parity: neighbour lists built by libjmd_b200.so vs the CPU oracle.

Bar: bit-exact.  With the default search grid the public `idx` arrays are
compared element by element (same neighbours in the same order as the
reference's candidate order); with the optional finer search grid the SETS are
compared (sorted rows / sorted pair lists, as BASELINE.json specifies).  Plus
shapes, occupancies, capacities, padding and error flags.
"""
import numpy as np
import pytest
import torch

from oracle import partition as opart
from oracle import space as ospace
from tests import util

pytestmark = pytest.mark.gpu

FORMATS = ['Dense', 'Sparse', 'OrderedSparse']


def _mods():
  import jax_md_b200 as jmd
  return jmd


def _dev(x):
  return torch.as_tensor(x, device='cuda')


def _build_both(R, box, r_cut, skin, fmt, periodic=True, **kw):
  jmd = _mods()
  if periodic:
    d_o, _ = ospace.periodic(box)
    d_g, _ = jmd.space.periodic(box)
  else:
    d_o, _ = ospace.free()
    d_g, _ = jmd.space.free()
  nf_o = opart.neighbor_list(d_o, box, r_cut, skin, format=opart.Format[fmt], **kw)
  nf_g = jmd.partition.neighbor_list(
      d_g, box, r_cut, skin, format=jmd.partition.NeighborListFormat[fmt], **kw)
  return nf_o, nf_g


def _assert_same(nb_o, nb_g, exact_order=True):
  assert nb_g.max_occupancy == nb_o.max_occupancy
  assert nb_g.cell_list_capacity == nb_o.cell_list_capacity
  assert int(nb_g.error.code) == int(nb_o.error)
  idx_g = nb_g.idx.cpu().numpy()
  assert idx_g.shape == nb_o.idx.shape
  assert idx_g.dtype == np.int32
  if exact_order:
    np.testing.assert_array_equal(idx_g, nb_o.idx)
  else:
    N = len(nb_o.reference_position)
    if nb_o.format is opart.Dense:
      np.testing.assert_array_equal(np.sort(idx_g, -1), np.sort(nb_o.idx, -1))
    else:
      np.testing.assert_array_equal(util.sparse_pairs(idx_g, N),
                                    util.sparse_pairs(nb_o.idx, N))
  np.testing.assert_array_equal(nb_g.reference_position.cpu().numpy(),
                                nb_o.reference_position)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dim', [2, 3])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_random_rectangular_box(fmt, dim, dtype):
  """reference tests/partition_test.py:203-301 configuration."""
  rng = np.random.default_rng(0)
  box = np.array([9.0, 4.0, 7.25][:dim], np.float32)
  N = 1000
  R = (rng.random((N, dim)) * box).astype(dtype)
  nf_o, nf_g = _build_both(R, box, 1.23, 0.0, fmt, capacity_multiplier=1.1)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  assert nb_o.use_cell_list
  _assert_same(nb_o, nb_g)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('brick', [0, 1, 2])
def test_fcc_lj_config(fmt, dtype, brick):
  """BASELINE config: LJ fcc, rho=0.8442, rc=2.5, skin 0.3 (N=4*12^3=6912);
  also with the cells stored brick-wise (same public idx, element by element)."""
  R, L = util.fcc(12, dtype=dtype)
  R = util.jitter(R, L, 0.05)
  nf_o, nf_g = _build_both(R, L, np.float32(2.5), np.float32(0.3), fmt)
  if brick:
    jmd = _mods()
    nf_g = jmd.partition.neighbor_list(
        jmd.space.periodic(L)[0], L, np.float32(2.5), np.float32(0.3),
        format=jmd.partition.NeighborListFormat[fmt], cell_brick_shift=brick)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  _assert_same(nb_o, nb_g)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('mask_self', [True, False])
def test_all_pairs_path(fmt, mask_self):
  """cutoff >= box/3 -> no cell list (partition.py:1052), candidates = all."""
  rng = np.random.default_rng(3)
  box = np.float32(5.0)
  R = (rng.random((64, 3)) * box).astype(np.float32)
  nf_o, nf_g = _build_both(R, box, 1.8, 0.2, fmt, mask_self=mask_self)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  assert not nb_o.use_cell_list
  _assert_same(nb_o, nb_g)


@pytest.mark.parametrize('case', [(0.12, True, 1.5), (0.25, False, 1.5),
                                  (0.31, False, 1.5), (0.31, False, 1.0)])
@pytest.mark.parametrize('mask_self', [False, True])
@pytest.mark.parametrize('fmt', FORMATS)
def test_issue191_capacity_goldens(case, mask_self, fmt):
  """reference tests/partition_test.py:488-546 (shape goldens)."""
  r_cut, disable, cm = case
  box = np.ones(3)
  R = np.ones((20, 3)) * 0.5
  want = {'Dense': (20, 19) if mask_self else (20, 20),
          'Sparse': (2, 380) if mask_self else (2, 400),
          'OrderedSparse': (2, 190)}[fmt]
  nf_o, nf_g = _build_both(R, box, r_cut, 0.1 * r_cut, fmt,
                           capacity_multiplier=cm, disable_cell_list=disable,
                           mask_self=mask_self)
  nb_g = nf_g.allocate(_dev(R))
  assert not bool(nb_g.did_buffer_overflow)
  assert tuple(nb_g.idx.shape) == want
  nb_o = nf_o.allocate(R)
  _assert_same(nb_o, nb_g)
  nb_g2 = nb_g.update(_dev(R + 0.1))
  assert not bool(nb_g2.did_buffer_overflow)
  assert tuple(nb_g2.idx.shape) == want
  nb_o2 = nb_o.update(R + 0.1)
  _assert_same(nb_o2, nb_g2)


def test_cell_list_overflow_flag():
  """reference tests/partition_test.py:363-401 (free space + cell list)."""
  R = np.array([[20., 20.], [30., 30.], [40., 40.], [50., 50.]], np.float32)
  nf_o, nf_g = _build_both(R, 100.0, 3.0, 0.0, 'Dense', periodic=False)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  _assert_same(nb_o, nb_g)
  R2 = np.array([[20., 20.], [20., 20.], [40., 40.], [50., 50.]], np.float32)
  nb_g = nb_g.update(_dev(R2))
  nb_o = nb_o.update(R2)
  assert bool(nb_g.did_buffer_overflow)
  assert int(nb_g.error.code) == int(nb_o.error)
  assert nb_g.idx.dtype == torch.int32


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('mode', ['fused', 'gated'])
def test_update_semantics(fmt, mode):
  """partition.py:1119-1154: no rebuild below the skin threshold, rebuild above
  it (strict >), sticky error bits; fused cooperative kernel and gated modes agree."""
  R, L = util.fcc(8, dtype=np.float32)
  R = util.jitter(R, L, 0.03)
  nf_o, nf_g = _build_both(R, L, np.float32(2.5), np.float32(0.3), fmt)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  nb_g._ws.update_mode = mode
  builds0 = nb_g._ws.state_host()[4]
  rng = np.random.default_rng(5)
  # (a) tiny move: below threshold -> identical list, no rebuild
  Ra = np.mod(R + rng.normal(0, 0.01, R.shape).astype(np.float32), L).astype(np.float32)
  nb_o = nb_o.update(Ra)
  nb_g = nb_g.update(_dev(Ra))
  assert not nb_o.did_rebuild
  assert nb_g._ws.state_host()[4] == builds0
  _assert_same(nb_o, nb_g)
  # (b) one atom jumps past skin/2 -> rebuild
  Rb = Ra.copy()
  Rb[17, 0] = np.mod(Rb[17, 0] + 0.2, L)
  nb_o = nb_o.update(Rb)
  nb_g = nb_g.update(_dev(Rb))
  assert nb_o.did_rebuild
  assert nb_g._ws.state_host()[4] == builds0 + 1
  _assert_same(nb_o, nb_g)
  # (c) crush a region: cell (and row) capacity overflow.  With a cell overflow
  # the reference loses atoms in an unspecified way, so only the cell bit and
  # the truthiness of did_buffer_overflow are defined; they must match + stick.
  Rc = Rb.copy()
  Rc[:200] = np.mod(Rb[100] + rng.normal(0, 0.4, (200, 3)).astype(np.float32), L)
  nb_o = nb_o.update(Rc)
  nb_g = nb_g.update(_dev(Rc))
  assert int(nb_g.error.code) & 2 == int(nb_o.error) & 2 == 2
  assert bool(nb_g.did_buffer_overflow) and bool(nb_o.did_buffer_overflow)
  nb_g = nb_g.update(_dev(Rb))
  assert int(nb_g.error.code) & 2 == 2                  # sticky


@pytest.mark.parametrize('fmt', FORMATS)
def test_neighbor_overflow_flag_all_pairs(fmt):
  """NEIGHBOR_LIST_OVERFLOW alone (no cell list involved): identical flag,
  identical truncated idx."""
  rng = np.random.default_rng(11)
  box = np.float32(6.0)
  R = (rng.random((96, 3)) * box).astype(np.float32)
  nf_o, nf_g = _build_both(R, box, 1.9, 0.2, fmt, capacity_multiplier=1.0)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  _assert_same(nb_o, nb_g)
  R2 = (R * np.float32(0.8)).astype(np.float32)        # denser -> more neighbours
  nb_o = nb_o.update(R2)
  nb_g = nb_g.update(_dev(R2))
  assert int(nb_o.error) == 1
  assert int(nb_g.error.code) == 1
  if fmt == 'Dense':
    np.testing.assert_array_equal(nb_g.idx.cpu().numpy(), nb_o.idx)


def test_large_random_sets_bit_exact():
  """SURVEY 7 hard-part 1 at scale: 100k random atoms, every row compared."""
  rng = np.random.default_rng(1)
  N = 100_000
  L = np.float32((N / 0.8442) ** (1 / 3))
  R = (rng.random((N, 3), np.float32) * L).astype(np.float32)
  nf_o, nf_g = _build_both(R, L, np.float32(2.5), np.float32(0.3), 'Dense')
  nb_g = nf_g.allocate(_dev(R))
  nb_o = nf_o.allocate(R)
  _assert_same(nb_o, nb_g)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dim', [2, 3])
def test_fine_search_grid_same_sets(fmt, dim):
  """Optional finer internal grid: identical neighbour SETS and flags."""
  rng = np.random.default_rng(4)
  box = np.array([19.0, 14.0, 17.25][:dim], np.float32)
  R = (rng.random((4000, dim)) * box).astype(np.float32)
  nf_o, nf_g = _build_both(R, box, 1.9, 0.2, fmt, fine_search_grid=True)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  assert nb_g._ws.c.stencil_w == 2
  _assert_same(nb_o, nb_g, exact_order=False)
  R2 = np.mod(R + rng.normal(0, 0.2, R.shape).astype(np.float32), box).astype(np.float32)
  _assert_same(nb_o.update(R2), nb_g.update(_dev(R2)), exact_order=False)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dim,dtype', [(3, np.float32), (3, np.float64), (2, np.float32)])
def test_warp_per_cell_scan_element_exact(fmt, dim, dtype):
  """The opt-in warp-per-cell candidate scan (`cell_scan=True`, csrc/jmd_nbr_cellscan.cuh):
  same lists as the oracle, element by element, on allocate and on update."""
  rng = np.random.default_rng(11)
  box = np.array([21.0, 16.5, 18.25][:dim], np.float32)
  n = 9000 if dim == 3 else 1500
  R = (rng.random((n, dim)) * box).astype(dtype)
  d_o, _ = ospace.periodic(box)
  nf_o = opart.neighbor_list(d_o, box, 2.1, 0.3, format=opart.Format[fmt])
  jmd = _mods()
  d_g, _ = jmd.space.periodic(box)
  nf_g = jmd.partition.neighbor_list(d_g, box, 2.1, 0.3, format=jmd.partition.NeighborListFormat[fmt],
                                     cell_scan=True)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  assert nb_g._ws.c.cell_scan == 1
  _assert_same(nb_o, nb_g)
  R2 = np.mod(R + rng.normal(0, 0.25, R.shape).astype(dtype), box).astype(dtype)
  _assert_same(nb_o.update(R2), nb_g.update(_dev(R2)))


def test_always_rebuild_when_skin_zero():
  R, L = util.fcc(6, dtype=np.float32)
  nf_o, nf_g = _build_both(R, L, 2.0, 0.0, 'Dense')
  nb_g = nf_g.allocate(_dev(R))
  b0 = nb_g._ws.state_host()[4]
  nb_g = nb_g.update(_dev(R))
  assert nb_g._ws.state_host()[4] == b0 + 1


def test_update_is_graph_capturable():
  """update() must not sync: capture it in a CUDA graph and replay."""
  R, L = util.fcc(8, dtype=np.float32)
  nf_o, nf_g = _build_both(R, L, np.float32(2.5), np.float32(0.3), 'Dense')
  Rd = _dev(R)
  nb_g = nf_g.allocate(Rd)
  nb_g._ws.update_mode = 'gated'
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(s):
    nb_g.update(Rd)
  torch.cuda.current_stream().wait_stream(s)
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    nb_g.update(Rd)
  b0 = nb_g._ws.state_host()[4]
  Rd.add_(0.2)          # every atom moved by > skin/2 -> replay must rebuild
  g.replay()
  torch.cuda.synchronize()
  assert nb_g._ws.state_host()[4] == b0 + 1


def _adversarial(N, L, cut, dtype, seed):
  """Random atoms plus partners placed within a few ulps of the list cutoff,
  atoms exactly on / just outside the box faces and unwrapped copies."""
  rng = np.random.default_rng(seed)
  R = (rng.random((N, 3)) * L).astype(dtype)
  m = N // 4
  u = rng.normal(size=(m, 3))
  u /= np.linalg.norm(u, axis=1, keepdims=True)
  eps = np.finfo(dtype).eps
  scale = 1.0 + rng.integers(-6, 7, (m, 1)) * eps * rng.choice([1, 4, 16, 64], (m, 1))
  R[m:2 * m] = (R[:m].astype(np.float64) + u * float(cut) * scale).astype(dtype)
  R[m:2 * m] = np.mod(R[m:2 * m], dtype(L))          # may round to exactly L: intended
  k = 2 * m
  R[k + 0] = [0, 0, 0]
  R[k + 1] = [L, L, L]                                # bins to cell 0, sits at L
  R[k + 2] = [-1e-6, 1.0, 2.0]
  R[k + 3] = [np.nextafter(dtype(L), dtype(0)), 0.5 * L, L]
  R[k + 4:k + 40] += dtype(L)                         # unwrapped: one box up
  R[k + 40:k + 80] -= dtype(3) * dtype(L)             # three boxes down
  return R


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_prefilter_adversarial_boundaries(fmt, dtype):
  """The stencil scan's contracted pre-filter must not change a single entry:
  pairs within a few ulps of cutoff^2, atoms on the faces, unwrapped atoms."""
  L = np.float32(21.7)
  cut, skin = np.float32(2.5), np.float32(0.3)
  R = _adversarial(4000, L, cut + skin, dtype, 7)
  nf_o, nf_g = _build_both(R, L, cut, skin, fmt)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  ws = nb_g._ws
  assert ws.c.no_filter == 0 and min(ws.c.fine_cps[k] for k in range(3)) >= 5
  _assert_same(nb_o, nb_g)
  R2 = _adversarial(4000, L, cut + skin, dtype, 8)
  # rebuild through update() (every atom moved): same kernels, gated launch
  _assert_same(nb_o.update(R2), nb_g.update(_dev(R2)))


@pytest.mark.parametrize('fmt', FORMATS)
def test_prefilter_equals_exact_scan_at_scale(fmt):
  """200k atoms (random + adversarial): default scan == exact_scan=True scan."""
  N = 200_000
  L = np.float32((N / 0.8442) ** (1 / 3))
  R = _adversarial(N, L, np.float32(2.8), np.float32, 9)
  jmd = _mods()
  d_g, _ = jmd.space.periodic(L)
  F = jmd.partition.NeighborListFormat[fmt]
  a = jmd.partition.neighbor_list(d_g, L, np.float32(2.5), np.float32(0.3), format=F).allocate(_dev(R))
  b = jmd.partition.neighbor_list(d_g, L, np.float32(2.5), np.float32(0.3), format=F,
                                  exact_scan=True).allocate(_dev(R))
  assert a._ws.c.no_filter == 0 and b._ws.c.no_filter == 1
  assert a.max_occupancy == b.max_occupancy
  assert torch.equal(a.idx, b.idx)
  assert int(a.error.code) == int(b.error.code)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_prefilter_2d_rectangular(fmt, dtype):
  """2-D rectangular box with >= 5 cells per side (pre-filter active): random
  atoms, near-cutoff partners and atoms on / outside the faces."""
  rng = np.random.default_rng(21)
  box = np.array([31.0, 24.5], np.float32)
  N, cut = 6000, 1.45
  R = (rng.random((N, 2)) * box).astype(dtype)
  m = N // 3
  ang = rng.random(m) * 2 * np.pi
  eps = np.finfo(dtype).eps
  scale = 1.0 + rng.integers(-5, 6, m) * eps * rng.choice([1, 8, 64], m)
  R[m:2 * m] = np.mod(R[:m].astype(np.float64) + (cut * scale)[:, None] *
                      np.stack([np.cos(ang), np.sin(ang)], 1), box.astype(np.float64)).astype(dtype)
  R[-1] = box
  R[-2] = [0, box[1]]
  R[-3] = [-1e-6, 3.0]
  R[-40:-3] += box.astype(dtype)                     # unwrapped
  nf_o, nf_g = _build_both(R, box, 1.2, 0.25, fmt)
  nb_o = nf_o.allocate(R)
  nb_g = nf_g.allocate(_dev(R))
  assert nb_g._ws.c.no_filter == 0 and min(nb_g._ws.c.fine_cps[k] for k in range(2)) >= 5
  _assert_same(nb_o, nb_g)


@pytest.mark.parametrize('fmt', FORMATS)
@pytest.mark.parametrize('mode', ['fused', 'gated'])
def test_lazy_idx_materialises_on_read(fmt, mode):
  """lazy_idx=True: update() leaves idx stale on the device after a rebuild and
  reading NeighborList.idx exports it; the array read is identical to the eager one."""
  jmd = _mods()
  R, L = util.fcc(12, dtype=np.float32)
  R = util.jitter(R, L, 0.05)
  d_g, _ = jmd.space.periodic(L)
  F = jmd.partition.NeighborListFormat[fmt]
  rng = np.random.default_rng(3)
  moves = [np.mod(R + rng.normal(0, s, R.shape).astype(np.float32), L).astype(np.float32)
           for s in (0.01, 0.2, 0.02, 0.25)]
  lists = []
  for lazy in (False, True):
    nf = jmd.partition.neighbor_list(d_g, L, np.float32(2.5), np.float32(0.3), format=F, lazy_idx=lazy)
    nb = nf.allocate(_dev(R))
    nb._ws.update_mode = mode
    seen = [nb.idx.clone()]
    for i, Rm in enumerate(moves):
      nb = nb.update(_dev(Rm))
      if lazy:
        pending = nb._ws.state_host()[8]
        assert pending == (1 if i in (1, 3) else 0) or i == 2   # rebuilds at the big moves
      if i != 1:                          # skip one read: the next rebuild supersedes it
        seen.append(nb.idx.clone())
        if lazy:
          assert nb._ws.state_host()[8] == 0
    lists.append((seen, nb.reference_position.clone(), int(nb.error.code)))
  for a, b in zip(lists[0][0], lists[1][0]):
    assert torch.equal(a, b)
  assert torch.equal(lists[0][1], lists[1][1]) and lists[0][2] == lists[1][2]
