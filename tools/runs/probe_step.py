"""Where does a step's wall time go?  CPU enqueue time vs GPU time, eager vs
CUDA-graph replay, tail-launch vs gated updates.  (development probe)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import jax_md_b200 as jmd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
R_h, box = bench.fcc((n, n, n)); N = len(R_h); L = box[0]
disp, shift = jmd.space.periodic(L)
nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, L, dr_threshold=bench.SKIN)
init_fn, apply_fn = jmd.simulate.nve(efn, shift, bench.DT)
Rd = torch.as_tensor(R_h, device='cuda'); Pd = torch.as_tensor(bench.momenta(N), device='cuda')
for mode in ('fused', 'gated'):
  nbrs = nf.allocate(Rd); nbrs._ws.update_mode = mode
  st = init_fn(0, Rd, kT=1.0, momenta=Pd, neighbor=nbrs)
  for _ in range(100):
    nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
  torch.cuda.synchronize()
  K = 300
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); e0.record()
  for _ in range(K):
    nbrs = nbrs.update(st.position); st = apply_fn(st, neighbor=nbrs)
  e1.record(); t1 = time.perf_counter()
  torch.cuda.synchronize(); t2 = time.perf_counter()
  print(f'{mode}: N={N} cpu enqueue {1e6*(t1-t0)/K:.1f} us/step, gpu {1e3*e0.elapsed_time(e1)/K:.1f} us/step, wall {1e6*(t2-t0)/K:.1f}')
# only the no-rebuild part: steps without update
torch.cuda.synchronize(); e0.record()
for _ in range(K):
  st = apply_fn(st, neighbor=nbrs)
e1.record(); torch.cuda.synchronize()
print(f'apply only (no update): gpu {1e3*e0.elapsed_time(e1)/K:.1f} us/step')
e0.record()
for _ in range(K):
  nbrs2 = nbrs.update(Rd)
e1.record(); torch.cuda.synchronize()
print(f'update only on build positions? gpu {1e3*e0.elapsed_time(e1)/K:.1f} us/step (rebuild each time if moved)')
