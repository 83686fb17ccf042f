"""Host side of space.periodic_general: the descriptor the kernels read (no GPU needed)."""
import numpy as np
import pytest
import torch

from jax_md_b200 import _lib, partition, space


def test_orthorhombic_forms_share_one_descriptor():
  L = np.array([10.0, 11.0, 12.0], np.float32)
  for box in (L, np.diag(L)):
    d, _ = space.periodic_general(box)
    st = space.space_struct(space.get_spec(d), 3, torch.float32)
    assert st.general == 1 and st.fractional == 1 and st.triclinic == 0
    np.testing.assert_array_equal(np.array(st.side[:]), L.astype(np.float64))
    np.testing.assert_array_equal(np.array(st.inv_box[:]), (np.float32(1) / L).astype(np.float64))


@pytest.mark.parametrize('dim', [2, 3])
def test_triclinic_descriptor(dim):
  H = np.array([[10.0, 2.5, -1.5], [0.0, 9.5, 2.0], [0.0, 0.0, 10.5]])[:dim, :dim]
  d, _ = space.periodic_general(H, fractional_coordinates=True)
  st = space.space_struct(space.get_spec(d), dim, torch.float64)
  assert st.general == 1 and st.triclinic == 1 and st.fractional == 1
  Hm = np.array(st.box_m[:]).reshape(3, 3)[:dim, :dim]
  Hi = np.array(st.inv_box_m[:]).reshape(3, 3)[:dim, :dim]
  np.testing.assert_array_equal(Hm, H)
  np.testing.assert_allclose(Hi @ H, np.eye(dim), atol=1e-14)
  # the bound the force kernels rely on: |d_j| <= half[j] for all j  =>  |(H^-1 d)_i| <= 1/2
  h = np.array(st.half[:dim])
  assert np.all(np.abs(Hi) @ h <= 0.5)
  rng = np.random.default_rng(0)
  dvec = (rng.random((1000, dim)) * 2 - 1) * h
  assert np.all(np.abs(dvec @ Hi.T) <= 0.5)


def test_triclinic_needs_unit_cube_positions_for_a_cell_grid():
  H = np.array([[10.0, 2.5, 0.0], [0.0, 9.5, 2.0], [0.0, 0.0, 10.5]])
  d, _ = space.periodic_general(H, fractional_coordinates=False)
  # partition.py:1052 `all(cell_size < box / 3)` sees the zero elements of a matrix box: the
  # reference never builds a cell grid for it, and neither does this host code
  nf = partition.neighbor_list(d, H, 2.0, 0.3)
  c = _lib.NbrT()
  use_cells, *_ = nf.allocate.fill_descriptor(c, 100, 3, np.float64, 100)
  assert not use_cells and c.use_cells == 0
  assert partition.is_box_valid(H) and not partition.is_box_valid(H.T)     # partition.py:676-681
  # partition.py:595-638: perpendicular widths of the cell, in f32
  cs = partition._fractional_cell_size(H.astype(np.float32), np.float32(2.3))
  assert cs == np.float32(0.25)
