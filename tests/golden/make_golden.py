"""Regenerates tests/golden/*.npz from the reference's own test DATA files.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The reference's Python cannot be imported here (no jax), so these fixtures are
the reference's committed golden vectors (data only, no source), re-packed so
they travel to the GPU box where /root/reference does not exist.

Sources (read-only):
  tests/data/simulation_test_state.npy     -> jammed_state.npz
      loader semantics: jax_md/test_util.py:347-358 (seven consecutive np.load)
      used by tests/quantity_test.py:134-150, tests/simulate_test.py:121-150
  tests/data/lammps_lj_stress_test{,_states} -> lammps_lj.npz
      parser semantics: jax_md/test_util.py:370-405; tests/quantity_test.py:436-455
  tests/data/lammps_npt_test               -> lammps_npt.npz
      parser semantics: jax_md/test_util.py:423-444 (columns 2:5 unit-cube positions, 5:8
      velocities, box 21.724 * I); the state of tests/simulate_test.py:586-690
Scalar goldens quoted in the reference tests are written to goldens.json.
"""
import json
import os

import numpy as np

REF = '/root/reference/tests/data'
OUT = os.path.dirname(os.path.abspath(__file__))


def jammed():
  with open(os.path.join(REF, 'simulation_test_state.npy'), 'rb') as f:
    names = ['fractional_position', 'real_position', 'species', 'sigma', 'box',
             'energy', 'pressure']
    d = {n: np.load(f) for n in names}
  np.savez(os.path.join(OUT, 'jammed_state.npz'), **d)


def lammps():
  with open(os.path.join(REF, 'lammps_lj_stress_test_states')) as f:
    data = f.read().split('\n')
  box = float(data[5].split(' ')[-1])
  R, V = [], []
  for l in data[9:-1]:
    R.append([float(x) for x in l.split(' ')[:3]])
    V.append([float(x) for x in l.split(' ')[3:]])
  with open(os.path.join(REF, 'lammps_lj_stress_test')) as f:
    row = [float(x) for x in f.read().split()]
  np.savez(os.path.join(OUT, 'lammps_lj.npz'), box=box, R=np.array(R),
           V=np.array(V), energy_per_atom=row[1], stress_row=np.array(row[2:]))


def lammps_npt():
  d = np.loadtxt(os.path.join(REF, 'lammps_npt_test'))
  np.savez_compressed(os.path.join(OUT, 'lammps_npt.npz'), box=np.eye(3) * 21.724,
                      position=d[:, 2:5], velocity=d[:, 5:8])


def scalars():
  g = {
      'sw_diamond_energy_per_atom': -4.336503155764325,   # tests/energy_test.py:429,464-466
      'sw_lattice_constant': 5.428,
      'jammed_energy': 0.45247561922261154,               # simulation_test_state.npy
      'jammed_pressure': 0.06307342050945483,
      'lammps_lj_energy_per_atom': -4.3523016,            # tests/data/lammps_lj_stress_test
      'issue191_shapes': {'Dense': [20, 19], 'Sparse': [2, 380],
                          'OrderedSparse': [2, 190]},     # tests/partition_test.py:516-546
  }
  with open(os.path.join(OUT, 'goldens.json'), 'w') as f:
    json.dump(g, f, indent=1)


if __name__ == '__main__':
  jammed()
  lammps()
  lammps_npt()
  scalars()
  print('wrote', sorted(os.listdir(OUT)))
