"""Eager loop vs lax.fori_loop (CUDA graph) step time at several sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
for n in (12, 20, 40, 63):
  R_h, box = bench.fcc((n, n, n)); N = len(R_h); L = box[0]
  disp, shift = jmd.space.periodic(L)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, L, dr_threshold=bench.SKIN, capacity_multiplier=1.5)
  init_fn, apply_fn = jmd.simulate.nve(efn, shift, bench.DT)
  Rd = torch.as_tensor(R_h, device='cuda'); Pd = torch.as_tensor(bench.momenta(N), device='cuda')
  nbrs = nf.allocate(Rd)
  st = init_fn(0, Rd, kT=1.0, momenta=Pd, neighbor=nbrs)
  def body(i, c):
    s, nb = c
    nb = nb.update(s.position)
    return apply_fn(s, neighbor=nb), nb
  for i in range(300): st, nbrs = body(i, (st, nbrs))
  torch.cuda.synchronize()
  K = 1000
  t0 = time.perf_counter()
  for i in range(K): st, nbrs = body(i, (st, nbrs))
  torch.cuda.synchronize(); te = (time.perf_counter() - t0) / K
  st, nbrs = jmd.lax.fori_loop(0, 100, body, (st, nbrs), unroll=50)
  g = jmd.lax.fori_loop.last
  torch.cuda.synchronize(); t0 = time.perf_counter()
  st, nbrs = jmd.lax.fori_loop(0, K, body, (st, nbrs), unroll=50, graph=g)
  torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / K
  print(f'N={N}: eager {1e6*te:.1f} us/step ({N/te:.3e} atom-steps/s), graph {1e6*tg:.1f} us/step ({N/tg:.3e}), overflow={bool(nbrs.did_buffer_overflow)}')
