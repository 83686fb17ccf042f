"""CPU-side cost of update() / apply() calls (development probe)."""
import sys, time, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd
n = 63
R_h, box = bench.fcc((n, n, n)); N = len(R_h); L = box[0]
disp, shift = jmd.space.periodic(L)
nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, L, dr_threshold=bench.SKIN)
init_fn, apply_fn = jmd.simulate.nve(efn, shift, bench.DT)
Rd = torch.as_tensor(R_h, device='cuda'); Pd = torch.as_tensor(bench.momenta(N), device='cuda')
nbrs = nf.allocate(Rd)
st = init_fn(0, Rd, kT=1.0, momenta=Pd, neighbor=nbrs)
for _ in range(50):
  nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
torch.cuda.synchronize()
K = 200
tu = ta = 0.0
for _ in range(K):
  t0 = time.perf_counter(); nbrs = nbrs.update(st.position); t1 = time.perf_counter()
  st = apply_fn(st, neighbor=nbrs); t2 = time.perf_counter()
  tu += t1 - t0; ta += t2 - t1
torch.cuda.synchronize()
print(f'cpu: update {1e6*tu/K:.1f} us, apply {1e6*ta/K:.1f} us per call')
pr = cProfile.Profile(); pr.enable()
for _ in range(K):
  nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
# bench-style timing with and without the NVML sampler thread
for sampler in (False, True):
  s = bench.ClockSampler(0) if sampler else None
  if s: s.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize(); e0.record()
  for _ in range(300):
    nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
  e1.record(); torch.cuda.synchronize()
  if s: s.stop()
  print('sampler', sampler, 'gpu us/step', 1e3 * e0.elapsed_time(e1) / 300)
