"""Timings of the other BASELINE.json configurations (parity-test cases, not the
bench.py headline): prints one JSON line per config.

  python benchmarks/configs.py [--quick]

 c1  examples/nve_neighbor_list.py as shipped: 2-D LJ, N=6400, f64, OrderedSparse
 c2  LJ fcc N=32,000 NVE, Dense and Sparse
 c3  bidisperse soft spheres N=256,000, 2-D and 3-D, FIRE
 c4  Stillinger-Weber diamond Si N=512,000, NVT Nose-Hoover
 f64 LJ fcc N=1,000,188 NVE in float64
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import jax_md_b200 as jmd  # noqa: E402


def timeit(step, state, nbrs, nf, steps, warmup):
  for _ in range(warmup):
    nbrs = nbrs.update(state.position)
    state = step(state, neighbor=nbrs)
  if bool(nbrs.did_buffer_overflow):
    nbrs = nf.allocate(state.position)
  torch.cuda.synchronize()
  b0 = nbrs._ws.state_host()[4]
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  def body(i, carry):
    st_, nb_ = carry
    nb_ = nb_.update(st_.position)
    return step(st_, neighbor=nb_), nb_
  # jit(lax.fori_loop) of the reference == CUDA-graph replay here; capture outside the timed region
  state, nbrs = jmd.lax.fori_loop(0, 20, body, (state, nbrs), unroll=20)
  g = jmd.lax.fori_loop.last
  steps = max(20, steps // 20 * 20)
  torch.cuda.synchronize()
  b0 = nbrs._ws.state_host()[4]
  e0.record()
  state, nbrs = jmd.lax.fori_loop(0, steps, body, (state, nbrs), unroll=20, graph=g)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / steps
  return ms, nbrs._ws.state_host()[4] - b0, bool(nbrs.did_buffer_overflow), state, nbrs


def report(name, N, ms, rebuilds, overflow, nbrs=None, itemsize=4, **extra):
  # whole-step HBM roofline fraction by the SURVEY 8(d) accounting: per step
  # sum_i n_i * (4 B index + 4 * itemsize B gathered position) + N * 6 * dim * itemsize B of state
  roof = None
  if nbrs is not None:
    ws = nbrs._ws
    pairs = int(torch.clamp(ws.t['cnt'][:N], max=ws.c.m_int).sum())
    step_bytes = pairs * (4 + 4 * itemsize) + N * 6 * ws.dim * itemsize
    peak, src = bench.peaks()
    roof = dict(pairs=pairs, step_bytes=step_bytes, achieved_gbs=step_bytes / (ms * 1e-3) / 1e9,
                peak_gbs=peak, frac=step_bytes / (ms * 1e-3) / 1e9 / peak, peak_source=src)
  print(json.dumps(dict(config=name, atoms=N, ms_per_step=ms, atom_steps_per_s=N / (ms * 1e-3),
                        rebuilds=int(rebuilds), overflow=overflow, roofline_step=roof, **extra)), flush=True)


def c1(steps, warmup):
  n = 80
  N = n * n
  L = np.float64(np.sqrt(N / 1.2) if False else n * 1.12)
  g = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing='ij'), -1).reshape(-1, 2) * 1.12
  R = torch.as_tensor(g.astype(np.float64), device='cuda')
  d, s = jmd.space.periodic(L)
  nf, efn = jmd.energy.lennard_jones_neighbor_list(d, L)          # OrderedSparse, skin 0.5
  nbrs = nf.allocate(R)
  init, step = jmd.simulate.nve(efn, s, 1e-3)
  st = init(0, R, kT=1e-3, neighbor=nbrs)
  ms, rb, ov, st, nbrs = timeit(step, st, nbrs, nf, steps, warmup)
  report('c1 2-D LJ N=6400 f64 OrderedSparse NVE (examples/nve_neighbor_list.py)', N, ms, rb, ov, nbrs, 8)


def c2(steps, warmup):
  R_h, box = bench.fcc((20, 20, 20))
  for fmt in ('Dense', 'Sparse'):
    d, s = jmd.space.periodic(box[0])
    nf, efn = jmd.energy.lennard_jones_neighbor_list(
        d, box[0], dr_threshold=0.3, format=jmd.partition.NeighborListFormat[fmt],
        capacity_multiplier=1.5)
    R = torch.as_tensor(R_h, device='cuda')
    nbrs = nf.allocate(R)
    init, step = jmd.simulate.nve(efn, s, 5e-3)
    st = init(0, R, kT=1.0, momenta=torch.as_tensor(bench.momenta(len(R_h)), device='cuda'), neighbor=nbrs)
    ms, rb, ov, st, nbrs = timeit(step, st, nbrs, nf, steps, warmup)
    report(f'c2 LJ fcc N=32000 f32 {fmt} NVE', len(R_h), ms, rb, ov, nbrs)


def c3(steps, warmup):
  N = 256_000
  for dim, dens in ((2, 0.8), (3, 0.9)):
    rng = np.random.default_rng(2)
    L = np.float32((N / dens) ** (1.0 / dim))
    R = torch.as_tensor((rng.random((N, dim)) * L).astype(np.float32), device='cuda')
    species = torch.as_tensor((np.arange(N) % 2).astype(np.int32), device='cuda')
    sigma = np.array([[1.0, 1.2], [1.2, 1.4]], np.float32)
    d, s = jmd.space.periodic(L)
    nf, efn = jmd.energy.soft_sphere_neighbor_list(d, L, species=species, sigma=sigma,
                                                   capacity_multiplier=1.5)
    nbrs = nf.allocate(R)
    init, step = jmd.minimize.fire_descent(efn, s)
    st = init(R, neighbor=nbrs)
    ms, rb, ov, st, nbrs = timeit(step, st, nbrs, nf, steps, warmup)
    report(f'c3 soft spheres {dim}-D N=256000 f32 OrderedSparse FIRE', N, ms, rb, ov, nbrs, 4,
           max_force=float(st.force.abs().max()))


def c4(steps, warmup):
  from tests import util
  R_h, L = util.diamond(40, a=5.431, dtype=np.float32)
  R = torch.as_tensor(R_h, device='cuda')
  d, s = jmd.space.periodic(np.float32(L))
  nf, efn = jmd.energy.stillinger_weber_neighbor_list(d, np.float32(L), capacity_multiplier=1.5)
  nbrs = nf.allocate(R)
  kT = 300 * 8.617333262e-5
  mass = 28.0855 * 1.03642698e-4          # metal units (eV ps^2 / A^2)
  dt = 1e-3
  init, step = jmd.simulate.nvt_nose_hoover(efn, s, dt, kT, chain_length=3, chain_steps=1,
                                            sy_steps=1, tau=100 * dt)
  st = init(0, R, mass=mass, neighbor=nbrs)
  ms, rb, ov, st, nbrs = timeit(step, st, nbrs, nf, steps, warmup)
  T = float(jmd.quantity.temperature(momentum=st.momentum, mass=st.mass)) / 8.617333262e-5
  report('c4 Stillinger-Weber Si N=512000 f32 Dense NVT Nose-Hoover', len(R_h), ms, rb, ov, nbrs, 4,
         temperature_K=T)


def f64(steps, warmup):
  R_h, box = bench.fcc((63, 63, 63), dtype=np.float64)
  d, s = jmd.space.periodic(box[0])
  nf, efn = jmd.energy.lennard_jones_neighbor_list(d, box[0], dr_threshold=0.3)
  R = torch.as_tensor(R_h, device='cuda')
  nbrs = nf.allocate(R)
  init, step = jmd.simulate.nve(efn, s, 5e-3)
  st = init(0, R, kT=1.0, momenta=torch.as_tensor(bench.momenta(len(R_h)).astype(np.float64), device='cuda'),
            neighbor=nbrs)
  ms, rb, ov, st, nbrs = timeit(step, st, nbrs, nf, steps, warmup)
  report('f64 LJ fcc N=1000188 f64 OrderedSparse NVE', len(R_h), ms, rb, ov, nbrs, 8)


if __name__ == '__main__':
  quick = '--quick' in sys.argv
  steps, warmup = (100, 50) if quick else (500, 200)
  for fn in (c1, c2, c3, c4, f64):
    fn(steps, warmup)
