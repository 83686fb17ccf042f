"""CPU oracle for the JAX MD short-range hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(`/root/reference/jax_md/{space,partition,smap,energy,simulate,minimize}.py`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; the product (`jax_md_b200`) never does
and fails loudly if its CUDA library is missing.

Parity status
-------------
* Energies / forces: PINNED against the reference's own golden values
  (tests/golden/, see tests/test_oracle_golden.py): jammed soft-sphere energy
  0.45247561922261154, Stillinger-Weber diamond -4.336503155764325 eV/atom,
  LAMMPS LJ E/N = -4.3523016, neighbour-list capacity goldens (20,19) / (2,380)
  / (2,190), cell placement goldens.
* Bit-exact neighbour sets / overflow flags at 1-ulp boundaries in f32:
  "parity unpinned".  `jax`/`jaxlib` are not importable in the build container
  (the arithmetic of the reference lives in that un-vendored dependency,
  pyproject.toml:44 `jax>=0.5.0`), and no reference test asserts set equality
  at the boundary.  The oracle uses NumPy's IEEE semantics (separately rounded
  multiply/add, `fmod`-based `mod`) for the op sequence the reference spells
  out (space.py:213-235).  XLA may contract FMAs differently.
"""
