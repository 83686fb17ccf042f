#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the short-range MD hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload: 3-D Lennard-Jones, f32, fcc at rho=0.8442, r_c=2.5,
skin 0.3, dt=0.005, kT=1.0, N = 4*n^3 atoms per GPU (n=63 -> 1,000,188).  One
"step" = `nbrs.update(R)` (skin predicate, rebuild when needed) + one
velocity-Verlet step through the public API.  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO, R_CUT, SKIN, DT, KT = 0.8442, 2.5, 0.3, 0.005, 1.0


def fcc(n_cells, dtype=np.float32):
  a = (4.0 / RHO) ** (1.0 / 3.0)
  nx, ny, nz = n_cells
  basis = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
  g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz),
                           indexing='ij'), -1).reshape(-1, 1, 3)
  R = ((g + basis[None]) * a).reshape(-1, 3).astype(dtype)
  box = np.array([nx * a, ny * a, nz * a], np.float32)
  return R, box


def momenta(N, seed=0):
  rng = np.random.default_rng(seed)
  p = rng.normal(0, np.sqrt(KT), (N, 3))
  return (p - p.mean(0, keepdims=True)).astype(np.float32)


class ClockSampler(threading.Thread):
  """Samples SM clocks / throttle reasons of one GPU through NVML while the
  timed region runs (the profiling recipe's clocks line)."""

  def __init__(self, index, period=0.05):
    super().__init__(daemon=True)
    self.index, self.period = index, period
    self.samples, self.reasons = [], set()
    self.max_mhz = None
    self._stop_evt = threading.Event()
    self.ok = False
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      # first queries of a process initialise NVML state: pay for that here, not in the timed region
      pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
      self.ok = True
    except Exception:   # pragma: no cover
      self.ok = False

  def run(self):
    if not self.ok:
      return
    nv = self.nv
    names = {
        getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
        getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
        getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
        getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
    }
    while not self._stop_evt.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if r & bit:
            self.reasons.add(name)
      except Exception:   # pragma: no cover
        pass
      time.sleep(self.period)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=2)
    med = float(np.median(self.samples)) if self.samples else None
    return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    with open(p) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------
# reference arm / cpu baseline: the NumPy oracle port on the host cores
# ------------------------------------------------------------------------------

def cpu_port_run(n, steps, warmup):
  """Times the CPU port of the reference path (oracle/c, multi-threaded C; the
  NumPy oracle is ~14x slower) on a bounded sample: fcc N = 4 n^3, `steps`
  update+NVE steps.  Returns (atom-steps/s, N, seconds, threads)."""
  from oracle import cport
  R, box = fcc((n, n, n))
  sysc = cport.LJSystem(R, box[0], r_cutoff=R_CUT, skin=SKIN, r_onset=2.0)
  sysc.run(momenta(len(R)), DT, warmup)
  t0 = time.perf_counter()
  sysc.run(None, DT, steps)
  dt = time.perf_counter() - t0
  th = sysc.threads
  sysc.close()
  return len(R) * steps / dt, len(R), dt, th


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  # the B200 arm's own single-GPU workload (N = 4 * cells^3 = 1,000,188 by default), its K and W
  # capped so that the run ends within minutes on the host cores; under torchrun (N GPUs) the
  # CPU arm still runs this one-GPU system on rank 0 -- a bounded sample of the N-GPU workload
  n = args.cpu_cells if args.cpu_cells > 0 else args.cells
  steps = max(1, min(args.steps, args.cpu_steps))
  warm = max(1, min(args.warmup, 5))
  v, N, secs, th = cpu_port_run(n, steps, warm)
  ms = secs / steps * 1e3
  sample = (f'LJ fcc N={N} (n={n}), {steps} update+NVE steps after {warm} warm-up steps through the C '
            f'port of the reference path on {th} threads (jax is not installable here: the reference '
            f'itself is not runnable)')
  line = {
      'impl': 'reference', 'metric': 'atom-timesteps/s', 'value': v,
      'unit': 'atom-timesteps/s', 'n_gpus': args.gpus, 'steps': steps,
      'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic',
      'config': {'workload': f'LJ fcc N={N} rho={RHO} rc={R_CUT} skin={SKIN} dt={DT} kT={KT} NVE, '
                             'C port of the reference path (CSR cell list, closed-form forces)',
                 'atoms': N},
      'cpu_baseline': {'value': v, 'unit': 'atom-timesteps/s', 'cores': th,
                       'kind': 'port', 'sample': sample},
      'e2e': {'value': v, 'unit': 'atom-timesteps/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------

def measure_single(cells, steps, warmup, fmt_name, unroll):
  """atom-steps/s of update + NVE on ONE GPU for an fcc box of `cells` (graph loop); the
  extra points of the JSON line (same physics and step as the headline)."""
  import torch
  import jax_md_b200 as jmd
  from jax_md_b200 import _lib
  R_h, box = fcc(cells)
  N = len(R_h)
  disp, shift = jmd.space.periodic(box)
  fmt = jmd.partition.NeighborListFormat[fmt_name]
  nf, efn = jmd.energy.lennard_jones_neighbor_list(disp, box, r_onset=2.0, r_cutoff=R_CUT,
                                                   dr_threshold=SKIN, format=fmt)
  init_fn, apply_fn = jmd.simulate.nve(efn, shift, DT)
  Rd = torch.as_tensor(R_h, device='cuda')
  nbrs = nf.allocate(Rd)
  state = init_fn(0, Rd, kT=KT, momenta=torch.as_tensor(momenta(N), device='cuda'), neighbor=nbrs)

  def body(i, carry):
    st, nb = carry
    nb = nb.update(st.position)
    return apply_fn(st, neighbor=nb), nb
  k = max(unroll, (max(warmup, 1) + unroll - 1) // unroll * unroll)
  state, nbrs = jmd.lax.fori_loop(0, k, body, (state, nbrs), unroll=unroll)
  g = jmd.lax.fori_loop.last
  if bool(nbrs.did_buffer_overflow):
    nbrs = nf.allocate(state.position)
    g = None
  torch.cuda.synchronize()
  b0 = nbrs._ws.state_host()[_lib.ST_BUILDS]
  k = max(unroll, steps // unroll * unroll)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  state, nbrs = jmd.lax.fori_loop(0, k, body, (state, nbrs), unroll=unroll, graph=g)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / k
  out = {'atoms': N, 'value': N / (ms * 1e-3), 'unit': 'atom-timesteps/s', 'ms_per_step': ms, 'steps': k,
         'rebuilds': int(nbrs._ws.state_host()[_lib.ST_BUILDS] - b0),
         'neighbor_overflow': bool(nbrs.did_buffer_overflow), 'n_gpus': 1}
  del nbrs, state, g
  jmd.lax.fori_loop.last = None
  torch.cuda.empty_cache()
  return out


def untimed_steps(args):
  """Steps every arm (1 GPU or N) runs before the timed window: the warm-up, one block of
  `unroll` steps in which the loop's CUDA graph is captured when the warm-up was too short to
  do it, and one more replayed block."""
  if args.loop != 'graph' or args.steps < args.unroll:
    return args.warmup
  return args.warmup + (args.unroll if args.warmup >= args.unroll else 2 * args.unroll)


def run_b200(args):
  import torch
  import torch.distributed as dist
  import jax_md_b200 as jmd
  from jax_md_b200 import _lib

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  n = args.cells
  if world > 1:
    from jax_md_b200 import domain
    return domain.bench_domain(args, world, rank, dev)

  R_h, box = fcc((n, n, n))
  N = len(R_h)
  P_h = momenta(N)
  L = box[0]
  disp, shift = jmd.space.periodic(L)
  fmt = jmd.partition.NeighborListFormat[args.format]
  nf, efn = jmd.energy.lennard_jones_neighbor_list(
      disp, L, r_onset=2.0, r_cutoff=R_CUT, dr_threshold=SKIN, format=fmt)
  init_fn, apply_fn = jmd.simulate.nve(efn, shift, DT)

  R_pin = torch.from_numpy(R_h).pin_memory()
  P_pin = torch.from_numpy(P_h).pin_memory()
  Rd = R_pin.to(dev, non_blocking=True)
  Pd = P_pin.to(dev, non_blocking=True)
  nbrs = nf.allocate(Rd)
  nbrs._ws.update_mode = args.update_mode
  state = init_fn(0, Rd, kT=KT, momenta=Pd, neighbor=nbrs)

  def make_body(apply):
    def body(i, carry):
      state, nbrs = carry
      nbrs = nbrs.update(state.position)
      state = apply(state, neighbor=nbrs)
      return state, nbrs
    return body

  body = make_body(apply_fn)
  loop = {'g': None}

  def md_steps(state, nbrs, k, body=body, loop=loop):
    # the loop of examples/nve_neighbor_list.py:186-195; --loop graph runs it through
    # jax_md_b200.lax.fori_loop (the jit(lax.fori_loop) of the reference: one CUDA
    # graph of `unroll` steps, replayed), --loop eager launches step by step
    if args.loop == 'graph' and k >= args.unroll:
      out = jmd.lax.fori_loop(0, k, body, (state, nbrs), unroll=args.unroll, graph=loop['g'])
      loop['g'] = jmd.lax.fori_loop.last
      return out
    for i in range(k):
      state, nbrs = body(i, (state, nbrs))
    return state, nbrs

  def barrier():
    torch.cuda.synchronize()

  # ---- warm-up (also melts the perfect lattice a little so rebuilds happen) ----
  state, nbrs = md_steps(state, nbrs, args.warmup)
  barrier()
  if bool(nbrs.did_buffer_overflow):
    nbrs = nf.allocate(state.position)
    nbrs._ws.update_mode = args.update_mode
    loop['g'] = None
  if args.loop == 'graph' and loop['g'] is None and args.steps >= args.unroll:
    # short warm-ups never reached the graph path: capture it now (one more untimed
    # block of steps) so that the timed region replays an existing graph
    state, nbrs = md_steps(state, nbrs, args.unroll)
    barrier()
  if args.loop == 'graph' and loop['g'] is not None:
    # one more untimed replay: the timed region then starts from a graph that has already been
    # replayed (the first replay after capture pays one-off upload / first-touch costs)
    state, nbrs = md_steps(state, nbrs, args.unroll)
    barrier()
  builds0 = nbrs._ws.state_host()[_lib.ST_BUILDS]

  # ---- timed region: exactly K steps, device timed ------------------------------
  sampler = ClockSampler(local)
  sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  state, nbrs = md_steps(state, nbrs, args.steps)
  e1.record()
  barrier()
  ms_total = e0.elapsed_time(e1)
  clocks = sampler.stop()
  builds = nbrs._ws.state_host()[_lib.ST_BUILDS] - builds0
  overflow = bool(nbrs.did_buffer_overflow)
  value = N * args.steps / (ms_total * 1e-3)

  # ---- variant: public idx materialised lazily (lazy_idx=True; not the headline) ---
  variants = {}
  if not args.no_variants:
    nf_l, efn_l = jmd.energy.lennard_jones_neighbor_list(
        disp, L, r_onset=2.0, r_cutoff=R_CUT, dr_threshold=SKIN, format=fmt, lazy_idx=True)
    init_l, apply_l = jmd.simulate.nve(efn_l, shift, DT)
    nb_l = nf_l.allocate(state.position)
    st_l = init_l(0, state.position.clone(), kT=KT, momenta=state.momentum.clone(), neighbor=nb_l)
    body_l, loop_l = make_body(apply_l), {'g': None}
    st_l, nb_l = md_steps(st_l, nb_l, 100, body_l, loop_l)
    barrier()
    b0 = nb_l._ws.state_host()[_lib.ST_BUILDS]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st_l, nb_l = md_steps(st_l, nb_l, args.steps, body_l, loop_l)
    b.record()
    barrier()
    ms_l = a.elapsed_time(b)
    idx_l = nb_l.idx                      # first read: the export runs now
    variants['lazy_idx'] = {
        'value': N * args.steps / (ms_l * 1e-3), 'ms_per_step': ms_l / args.steps,
        'rebuilds': int(nb_l._ws.state_host()[_lib.ST_BUILDS] - b0),
        'note': 'lazy_idx=True: rebuilds inside update() leave the public idx stale on the device; '
                'it is exported when NeighborList.idx is read (once, after the loop, here). The '
                'headline value keeps the default: idx rewritten at every rebuild.',
        'idx_shape_after_read': list(idx_l.shape)}
    del nb_l, st_l, idx_l, loop_l

  # ---- neighbour rebuild time (one full forced rebuild) -------------------------
  ws = nbrs._ws
  pp = _lib.ptr(state.position)
  rb = []
  for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = _lib.stream()
    _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, st)
    _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 0, st)
    _lib.call('jmd_nbr_export', ws.ref(), pp, 0, st)
    b.record()
    torch.cuda.synchronize()
    rb.append(a.elapsed_time(b))
  rebuild_ms = float(np.median(rb))

  # ---- roofline of the dominant kernel (fused force + half kick), timed alone ---
  pairs = int(torch.minimum(ws.t['cnt'][:N], torch.tensor(ws.c.m_int, device=dev)).sum())
  kernel_bytes = pairs * 20 + N * 60       # DESIGN.md "algorithmic bytes"
  step_bytes = pairs * 20 + N * 72
  fstep = apply_fn._stepper
  kt = []
  P_tmp = state.momentum.clone()
  kw = {}
  _, species, params = efn._resolve(nbrs, kw)
  for i in range(args.kernel_reps + 3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    efn.launch(state.position, nbrs, species, params, want_energy=False,
               momentum=P_tmp, mass=state.mass, dt_2=0.0, red=fstep.red(state.position),
               refresh_positions=False)
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
      kt.append(a.elapsed_time(b))
  k_ms = float(np.mean(kt))
  peak, peak_src = peaks()
  achieved = kernel_bytes / (k_ms * 1e-3) / 1e9
  roofline = {'bound': 'hbm', 'kernel': 'k_pair_force<float,3,LJ,scalar,kick>',
              'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
              'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
              'kernel_ms': k_ms, 'algorithmic_bytes_per_launch': kernel_bytes,
              'step_frac': step_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}

  # ---- end to end through the public API with host buffers ----------------------
  blk = args.block
  n_blocks = max(1, args.steps // blk)
  out_R = torch.empty_like(R_pin).pin_memory()
  out_P = torch.empty_like(P_pin).pin_memory()
  flags = torch.zeros(2, dtype=torch.float64).pin_memory()
  if args.loop == 'graph' and loop['g'] is None and blk >= args.unroll:
    state, nbrs = md_steps(state, nbrs, args.unroll)     # capture outside the timed region
  R_pin.copy_(state.position.cpu())
  P_pin.copy_(state.momentum.cpu())
  barrier()
  # The neighbour list is allocated and the loop's CUDA graph captured (above): the
  # counterpart of the reference's first allocate + jit compile, which SURVEY 8(d)
  # keeps out of the metric.  Timed: H2D of the state from pinned host memory,
  # update() of the existing list on it, init (first force evaluation), the steps with
  # the periodic host readbacks of the example loop, D2H of the final state.
  t0 = time.perf_counter()
  Rd = R_pin.to(dev, non_blocking=True)
  Pd = P_pin.to(dev, non_blocking=True)
  nb2 = nbrs.update(Rd)
  st2 = init_fn(0, Rd, kT=KT, momenta=Pd, neighbor=nb2)
  d2h = 0
  for _ in range(n_blocks):
    st2, nb2 = md_steps(st2, nb2, blk)
    flags[0] = float(nb2.did_buffer_overflow)          # sync, like the example loop
    flags[1] = float(fstep.red(st2.position)[_lib.RED_KINETIC])
    d2h += 1 + 8
  out_R.copy_(st2.position, non_blocking=True)
  out_P.copy_(st2.momentum, non_blocking=True)
  barrier()
  e2e_s = time.perf_counter() - t0
  e2e_steps = n_blocks * blk
  h2d = R_pin.numel() * 4 + P_pin.numel() * 4
  d2h += out_R.numel() * 4 + out_P.numel() * 4
  e2e = {'value': N * e2e_steps / e2e_s, 'unit': 'atom-timesteps/s',
         'h2d_bytes_per_step': h2d / e2e_steps, 'd2h_bytes_per_step': d2h / e2e_steps,
         'steps': e2e_steps, 'includes': 'H2D of state from pinned memory, update() of the '
         f'allocated list, init (forces), {e2e_steps} update+apply steps, overflow/KE readback '
         f'every {blk} steps, D2H of state; list allocation and graph capture (first allocate / '
         'jit compile of the reference) happen before the timed region'}

  # ---- extra points (not the headline): the per-GPU load of the north star's N=32M run
  # (4M atoms) and the 8M-atom system of the strong-scaling series, on this one GPU
  extra = {}
  if not args.no_extra:
    del nbrs, state, st2, nb2
    loop['g'] = None
    torch.cuda.empty_cache()
    for key, cells in (('weak_4m_per_gpu', (100, 100, 100)), ('strong_8m_total', (8 * n, n, n))):
      extra[key] = measure_single(cells, args.steps, args.warmup, args.format, args.unroll)

  # ---- CPU baseline (oracle port, bounded sample) -------------------------------
  cpu = None
  if not args.no_cpu:
    v, Nc, secs, th = cpu_port_run(args.cpu_cells if args.cpu_cells > 0 else args.cells, args.cpu_steps, 2)
    cpu = {'value': v, 'unit': 'atom-timesteps/s', 'cores': th, 'kind': 'port',
           'sample': f'LJ fcc N={Nc}, {args.cpu_steps} update+NVE steps, C port of the '
                     f'reference path on {th} threads ({secs:.1f} s)'}

  # kernels launched per step (all of them ours; gated launches that find
  # nothing to do on a non-rebuild step are still launches): fused update =
  # k_update + k_nbr_stencil_scan + k_update_c, then k_kick_drift + k_pair_force
  if args.update_mode == 'fused':
    # k_update, stencil scan (gated), [look-back offsets], export (gated), drift, force
    per_step = 5 if args.format == 'Dense' else 6
  else:
    per_step = 1 + 8 + 2 + (2 if args.format == 'Dense' else 7) + 2
  # DRAM bytes per launch of this kernel: not measurable inside the run (needs ncu); taken
  # from the committed ncu --set full capture of the same workload when there is one
  # (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep), else null
  roofline['traffic'] = None
  tp = os.path.join(ROOT, 'profiles', 'traffic.json')
  if os.path.exists(tp):
    with open(tp) as f:
      rec = json.load(f).get('k_pair_force')
    if rec and rec.get('atoms') == N and rec.get('format') == args.format:
      roofline['traffic'] = rec['dram_bytes_read'] + rec['dram_bytes_write']
      roofline['traffic_source'] = rec['source']
  line = {
      'metric': 'atom-timesteps/s', 'value': value, 'unit': 'atom-timesteps/s',
      'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'LJ fcc N={N} rho={RHO} rc={R_CUT} skin={SKIN} dt={DT} '
                             f'kT={KT} NVE, neighbour format {args.format}',
                 'atoms': N, 'l2_policy': 'working set (idx rows %.0f MB + state) exceeds L2'
                 % (pairs * 4 / 1e6), 'rebuilds_in_timed_region': int(builds),
                 'untimed_steps_before_window': untimed_steps(args),
                 'loop': ('jax_md_b200.lax.fori_loop: CUDA graph of %d steps, replayed' % args.unroll)
                 if args.loop == 'graph' else 'eager Python loop',
                 'neighbor_overflow': overflow},
      'neighbor_rebuild_ms': rebuild_ms,
      'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'variants': variants,
      'weak_4m_per_gpu': extra.get('weak_4m_per_gpu'), 'strong_8m_total': extra.get('strong_8m_total'),
      'gpu_launches': int(args.steps * per_step),
      'clocks': clocks,
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=1000)
  ap.add_argument('--warmup', type=int, default=500)
  ap.add_argument('--impl', default='b200')
  ap.add_argument('--cells', type=int, default=63, help='fcc cells per side per GPU')
  ap.add_argument('--format', default='OrderedSparse',
                  choices=['Dense', 'Sparse', 'OrderedSparse'])
  ap.add_argument('--block', type=int, default=100)
  ap.add_argument('--kernel-reps', type=int, default=20)
  ap.add_argument('--cpu-cells', type=int, default=0,
                  help='fcc cells per side of the CPU arms (0: the same system as the GPU arm)')
  ap.add_argument('--cpu-steps', type=int, default=30)
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-variants', action='store_true')
  ap.add_argument('--no-extra', action='store_true', help='skip the 4M / 8M-atom extra points')
  ap.add_argument('--loop', default='graph', choices=['eager', 'graph'])
  ap.add_argument('--unroll', type=int, default=20, help='steps per captured CUDA graph')
  ap.add_argument('--update-mode', default='fused', choices=['fused', 'gated'],
                  help="how update()'s lax.cond is realised: one cooperative kernel or gated kernels")
  args = ap.parse_args()
  if args.impl == 'reference':
    return run_reference(args)
  run_b200(args)


if __name__ == '__main__':
  main()
