import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
for n in (12, 40):
  R_h, box = bench.fcc((n, n, n)); L = float(box[0])
  rng = np.random.default_rng(0)
  S = np.mod(R_h / L + rng.normal(0, 0.002, R_h.shape), 1.0).astype(np.float32)
  Sd = torch.as_tensor(S, device='cuda')
  d, s = jmd.space.periodic_general(np.float32(L))
  nf, efn = jmd.energy.lennard_jones_neighbor_list(d, np.float32(L), r_onset=2.0, r_cutoff=2.5, dr_threshold=0.3,
                                                   fractional_coordinates=True, format=jmd.partition.Dense,
                                                   capacity_multiplier=1.5)
  nb = nf.allocate(Sd)
  init, step = jmd.simulate.npt_nose_hoover(efn, s, 2e-3, 1.0, 1.0)
  st = init(0, Sd, np.float32(L), neighbor=nb)
  print('N', len(S), 'init dUdV', float(st.dUdV), 'box_mass', float(st.box_mass), 'KE', float(jmd.simulate.kinetic_energy(st)),
        'thermo', st.thermostat._buf.tolist(), 'baro', st.barostat._buf.tolist(), flush=True)
  for i in range(4):
    try:
      nb = nb.update(st.position, box=jmd.simulate.npt_box(st)); st = step(st, neighbor=nb)
    except Exception as e:
      print('step', i, 'FAILED', repr(e)[:200]); break
    print('step', i, 'box_pos', float(st.box_position), 'box_mom', float(st.box_momentum), 'dUdV', float(st.dUdV),
          'KE', float(jmd.simulate.kinetic_energy(st)), 'box00', float(jmd.simulate.npt_box(st)[0, 0]), 'ovf', bool(nb.did_buffer_overflow), flush=True)
