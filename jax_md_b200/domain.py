"""Slab domain decomposition of the short-range MD step over the GPUs of one
node.  One process per GPU.

The reference has no API for this (its only multi-device code is the TPU voxel
prototype, jax_md/tpu.py:481-525,1498-1555), so this module is additive.  The
physics is the single-GPU path unchanged: the local system is
`owned atoms + ghost atoms` in GLOBAL coordinates with the GLOBAL periodic box,
so the same neighbour-list and force kernels run on it; rows, forces and skin
checks exist for owned atoms only (`jmd_nbr_t.n_rows`), hence no reverse
(force) communication.

A step is FOUR kernels and no host work besides one CUDA-graph launch:
  jmd_nve_kick_drift   half kick + drift of the owned atoms; leaves the skin flags
  jmd_dd_comm_push     gathers the face atoms and stores them straight into the ring
                       neighbours' landing rows through CUDA-IPC peer mappings (NVLink),
                       releases their signal words; broadcasts this rank's rebuild flag
  jmd_dd_comm_wait     waits for both neighbours' rows, copies them behind the owned atoms
                       and into the cell-sorted float4 array; ORs all ranks' flags and
                       publishes the global decision to a mapped host word
  jmd_pair_force       forces on the owned rows + second half kick
All counts the kernels need live on the device (`jmd_nbr_t.n_dev`, `info`), so the
captured graph stays valid across rebuilds.  The host polls the decision word while the
force kernel is still running; only a rebuild (every ~8 steps) is host-driven:
atom migration and ghost re-selection (ordered select kernels, fixed-capacity messages
that carry their counts, `torch.distributed` send/recv), one host read of the new
counts, then the local neighbour-list rebuild.
"""
import contextlib
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, partition, smap, space

f32 = np.float32
_INT_MAX = 2**31 - 1


class RingComm:
  """Periodic ring of ranks along the decomposition axis.  Device agnostic:
  with the NCCL backend tensors go straight over NVLink, otherwise (gloo) they
  are staged through host memory, which lets CPU and single-GPU tests drive the
  same code."""

  def __init__(self, group=None):
    self.group = group
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.left = (self.rank - 1) % self.world
    self.right = (self.rank + 1) % self.world
    self.direct = dist.is_initialized() and dist.get_backend(group) == 'nccl'
    # the per-step rebuild flag is all-reduced on its own communicator (own NCCL
    # stream), so it never queues in front of the halo exchange
    self.flag_group = dist.new_group(backend='nccl') if (self.direct and self.world > 1) else group

  def _stage(self, t):
    return t if (self.direct or not t.is_cuda) else t.cpu()

  def exchange(self, send_left, send_right, recv_left, recv_right):
    """send_left -> left neighbour (arrives there as its `recv_right`),
    send_right -> right neighbour.  All tensors contiguous; sizes agreed
    beforehand (exchange_counts) or fixed (messages that carry their count)."""
    self.exchange_many([(send_left, send_right, recv_left, recv_right)])

  def exchange_many(self, messages):
    """Several (send_left, send_right, recv_left, recv_right) groups in ONE
    batched send/recv (one NCCL group launch)."""
    if self.world == 1:
      for sl, sr, rl, rr in messages:
        rr.copy_(sl)
        rl.copy_(sr)
      return
    ops, staged = [], []
    for send_left, send_right, recv_left, recv_right in messages:
      sl, sr = self._stage(send_left), self._stage(send_right)
      rl = recv_left if (self.direct or not recv_left.is_cuda) else torch.empty_like(recv_left, device='cpu')
      rr = recv_right if (self.direct or not recv_right.is_cuda) else torch.empty_like(recv_right, device='cpu')
      # With two ranks left == right: messages between one pair match in issue
      # order, so receive-from-right is posted before receive-from-left.
      if sl.numel():
        ops.append(dist.P2POp(dist.isend, sl, self.left, self.group))
      if sr.numel():
        ops.append(dist.P2POp(dist.isend, sr, self.right, self.group))
      if rr.numel():
        ops.append(dist.P2POp(dist.irecv, rr, self.right, self.group))
      if rl.numel():
        ops.append(dist.P2POp(dist.irecv, rl, self.left, self.group))
      staged.append((rl, recv_left, rr, recv_right))
    if ops:
      for w in dist.batch_isend_irecv(ops):
        w.wait()
    for rl, recv_left, rr, recv_right in staged:
      if rl is not recv_left:
        recv_left.copy_(rl)
      if rr is not recv_right:
        recv_right.copy_(rr)

  def exchange_counts(self, n_to_left, n_to_right):
    """-> (n_from_left, n_from_right) as Python ints (host sync)."""
    if self.world == 1:
      return int(n_to_right), int(n_to_left)
    dev = 'cuda' if self.direct else 'cpu'
    s_l = torch.tensor([int(n_to_left)], dtype=torch.int64, device=dev)
    s_r = torch.tensor([int(n_to_right)], dtype=torch.int64, device=dev)
    r_l = torch.zeros(1, dtype=torch.int64, device=dev)
    r_r = torch.zeros(1, dtype=torch.int64, device=dev)
    self.exchange(s_l, s_r, r_l, r_r)
    return int(r_l.item()), int(r_r.item())

  def any(self, flag_tensor):
    """Global OR of a 1-element integer tensor -> Python bool (host sync)."""
    if self.world > 1:
      t = flag_tensor if (self.direct or not flag_tensor.is_cuda) else flag_tensor.cpu()
      dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
      return bool(t.item() != 0)
    return bool(flag_tensor.item() != 0)

  def max_int(self, value):
    """Global maximum of a Python int (setup time; host sync)."""
    if self.world == 1:
      return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device='cuda' if self.direct else 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
    return int(t.item())

  def sum(self, t):
    if self.world > 1:
      if self.direct or not t.is_cuda:
        dist.all_reduce(t, group=self.group)
      else:
        h = t.cpu()
        dist.all_reduce(h, group=self.group)
        t.copy_(h)
    return t


class SlabState:
  """Local part of the system: capacity-sized arrays, owned atoms first."""

  def __init__(self, R, P, F, gid, n_own):
    self.R, self.P, self.F, self.gid = R, P, F, gid
    self.n_own, self.n_ghost = n_own, 0

  @property
  def position(self):
    return self.R[:self.n_own]

  @property
  def momentum(self):
    return self.P[:self.n_own]

  @property
  def force(self):
    return self.F[:self.n_own]

  @property
  def global_id(self):
    return self.gid[:self.n_own]


class SlabDomain:
  """NVE over a slab decomposition along `axis` for a fused pair energy
  function (`smap.PairNeighborListFn`) built with the GLOBAL periodic space."""

  def __init__(self, box, energy_fn, r_cutoff, dr_threshold, dt, comm=None,
               axis=0, mass=1.0, capacity_factor=1.3, capacity_multiplier=1.25,
               transport=None, use_graph=True):
    _lib.require_cuda()
    self.comm = comm or RingComm()
    self.box = np.asarray(box, np.float64).reshape(-1)
    self.dim = len(self.box)
    self.axis = axis
    self.energy_fn = energy_fn
    self.r_cutoff, self.skin = float(r_cutoff), float(dr_threshold)
    self.dt = float(f32(dt))
    self.dt_2 = float(f32(f32(dt) / 2))
    self.mass_value = float(mass)
    self.capacity_factor = capacity_factor
    self.capacity_multiplier = capacity_multiplier
    W, r = self.comm.world, self.comm.rank
    self.width = self.box[axis] / W
    self.lo = r * self.width
    # ghosts: everything within (cutoff + margin) of a face; the margin makes it
    # a strict superset of what the exact-arithmetic candidate test can accept
    self.ghost_width = (self.r_cutoff + self.skin) * (1.0 + 1e-4) + 1e-6
    if W > 1 and self.width < 2 * self.ghost_width:
      raise ValueError('slab narrower than two ghost layers')
    self.disp, self.shift = space.periodic(self.box.astype(np.float32) if self.dim > 1 else f32(self.box[0]))
    self.neighbor_fn = partition.neighbor_list(
        self.disp, self.box.astype(np.float32), f32(r_cutoff), f32(dr_threshold),
        capacity_multiplier=capacity_multiplier, format=partition.Dense)
    self.nbrs = None
    self.rebuilds = 0
    self._lists = None
    # the halo carries positions only: per-atom species / parameters would need their own
    # exchange: refuse them instead of silently treating ghosts as species 0
    if getattr(energy_fn, 'species', None) is not None:
      raise NotImplementedError('SlabDomain: energy functions with species are not supported')
    for k, v in getattr(energy_fn, 'kwargs', {}).items():
      if isinstance(v, (torch.Tensor, np.ndarray)) and getattr(v, 'ndim', 0) > 0:
        raise NotImplementedError(f'SlabDomain: per-atom / table parameter {k!r} is not supported')
    self.transport = transport or os.environ.get('JMD_DD_TRANSPORT', 'p2p')
    self.use_graph = use_graph and os.environ.get('JMD_DD_GRAPH', '1') != '0'
    self._graph = None
    self._epoch = 0
    self._pending = None
    self._peer = None
    self._warm = 0

  # -- setup ------------------------------------------------------------------------
  def init(self, R_own, P_own, gid_own=None):
    """R_own / P_own: this rank's atoms (any order), CUDA tensors [n, dim]."""
    dev, dt = R_own.device, R_own.dtype
    n = R_own.shape[0]
    # capacities are sized from the largest slab so that the fixed-size rebuild
    # messages have the same length on every rank
    n_max = self.comm.max_int(n)
    cap = int(n_max * self.capacity_factor) + 1024
    n_ghost_est = 0
    if self.comm.world > 1:
      n_ghost_est = int(2 * n_max * self.ghost_width / self.width * self.capacity_factor) + 1024
    self.cap = cap + n_ghost_est
    self.cap_list = max(n_ghost_est, int(0.25 * n_max) + 1024)
    R = torch.zeros((self.cap, self.dim), dtype=dt, device=dev)
    P = torch.zeros_like(R)
    F = torch.zeros_like(R)
    gid = torch.full((self.cap,), -1, dtype=torch.int64, device=dev)
    R[:n] = R_own
    P[:n] = P_own
    gid[:n] = gid_own if gid_own is not None else torch.arange(n, device=dev)
    st = SlabState(R, P, F, gid, n)
    self.dtype, self.device = dt, dev
    self.dtc = _lib.dtype_code(dt)
    self.mass = torch.full((1,), self.mass_value, dtype=dt, device=dev)
    self.red = torch.zeros(_lib.RED_COUNT, dtype=torch.float64, device=dev)
    self.partials = smap.Scratch.get(self.cap, dev)
    self.sp = space.space_struct(space.get_spec(self.shift), self.dim, dt)
    # Rebuild-time pipeline (no host round trip before the single read of `info`):
    # fixed-capacity selection lists and messages that carry their own counts.
    i32 = dict(dtype=torch.int32, device=dev)
    self.cap_mig = int(0.02 * n_max) + 4096
    self.list_a = torch.empty(self.cap_list, **i32)       # face atoms (sorted)
    self.list_b = torch.empty(self.cap_list, **i32)
    self.mig_a = torch.empty(self.cap_mig, **i32)         # leavers (sorted)
    self.mig_b = torch.empty(self.cap_mig, **i32)
    self.counters = torch.zeros(2, **i32)                 # face counts
    self.mig_counters = torch.zeros(2, **i32)
    self.info = torch.zeros(_lib.DD_INFO_COUNT, **i32)
    self.info_host = torch.zeros(_lib.DD_INFO_COUNT, dtype=torch.int32).pin_memory()
    self.mig_scratch = torch.empty(6 * self.cap_mig, **i32)
    self.sel_scratch = torch.zeros(self.cap // 2048 + 4, dtype=torch.int64, device=dev)
    t = dict(dtype=dt, device=dev)
    self.mig_pay = [torch.zeros((self.cap_mig, 3 * self.dim), **t) for _ in range(4)]     # out l, r; in l, r
    self.mig_gid = [torch.zeros(self.cap_mig + 1, dtype=torch.int64, device=dev) for _ in range(4)]
    self.face_msg = [torch.zeros((self.cap_list + 1, self.dim), **t) for _ in range(4)]   # out l, r; in l, r
    self._send_l = torch.empty((self.cap_list, self.dim), **t)
    self._send_r = torch.empty((self.cap_list, self.dim), **t)
    self.info[_lib.DD_N_OWN] = n
    self._flag_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    self._flag_host = torch.zeros(1, dtype=torch.int64).pin_memory()
    self._flag_event = torch.cuda.Event()
    self._side_stream = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None
    self._decision_pending = False
    self._rebuild(st, first=True)
    if self.transport == 'p2p':
      self._setup_peers(st)
    self._force(st, kick=False)
    return st

  def reset(self, st, R_own, P_own, gid_own=None):
    """Loads a new local state into the initialised domain (buffers, peer mappings and
    the captured step graph are kept): rebuild + first force evaluation."""
    n = R_own.shape[0]
    if n > self.cap:
      raise ValueError('state larger than the slab capacity')
    st.R[:n] = R_own
    st.P[:n] = P_own
    st.gid[:n] = gid_own if gid_own is not None else torch.arange(n, device=st.R.device)
    st.n_own, st.n_ghost = n, 0
    self.info.zero_()
    self.info[_lib.DD_N_OWN] = n
    if self._peer is not None and self._epoch > 0 and self._pending is None:
      self._poll(self._epoch)                 # the last step's decision is moot: rebuild anyway
    self._pending = None
    self._rebuild(st)
    self._force(st, kick=False)
    return st

  # -- peer-memory exchange (CUDA IPC) --------------------------------------------------
  def _setup_peers(self, st):
    """One shared block per rank: landing rows [2 parities][2 sides][cap_list][dim], two
    signal words, one flag word per rank.  Handles go round with all_gather_object; the
    neighbours' blocks are mapped with cudaIpcOpenMemHandle (NVLink peer access)."""
    W, r = self.comm.world, self.comm.rank
    item = 4 if self.dtype == torch.float32 else 8
    land_bytes = 2 * 2 * self.cap_list * self.dim * item
    land_bytes = (land_bytes + 255) // 256 * 256
    total = land_bytes + 256 + 2 * 8 * max(W, 32)
    ptr = C.c_void_p()
    handle = (C.c_uint8 * 64)()
    _lib.call('jmd_p2p_alloc', total, C.byref(ptr), handle)
    base = ptr.value
    handles = [None] * W
    if W > 1:
      dist.all_gather_object(handles, bytes(handle), group=self.comm.group)
    else:
      handles[0] = bytes(handle)
    mapped = {r: base}

    def peer(rank):
      if rank not in mapped:
        q = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handles[rank])
        _lib.call('jmd_p2p_open', buf, C.byref(q))
        mapped[rank] = q.value
      return mapped[rank]
    flags_ptrs = torch.tensor([peer(k) + land_bytes + 256 for k in range(W)], dtype=torch.int64,
                              device=self.device)
    host = C.c_void_p()
    hdev = C.c_void_p()
    _lib.call('jmd_host_flag_alloc', C.byref(host), C.byref(hdev))
    self._host_flag = C.c_uint64.from_address(host.value)
    ws = self.nbrs._ws
    dd = _lib.DdT()
    dd.dtype, dd.dim, dd.rank, dd.world = self.dtc, self.dim, r, W
    dd.cap_list = self.cap_list
    dd.always_rebuild = int(ws.c.always_rebuild)
    dd.face_l, dd.face_r = self.list_a.data_ptr(), self.list_b.data_ptr()
    dd.face_counts = self.counters.data_ptr()
    dd.info = self.info.data_ptr()
    self._epoch_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
    self._ticket = torch.zeros(2, dtype=torch.int32, device=self.device)
    dd.epoch, dd.ticket = self._epoch_dev.data_ptr(), self._ticket.data_ptr()
    dd.skin_blk = ws.t['skin_blk'].data_ptr()
    dd.land, dd.signal, dd.flags = base, base + land_bytes, base + land_bytes + 256
    left, right = peer(self.comm.left), peer(self.comm.right)
    dd.peer_land_l, dd.peer_land_r = left, right
    dd.peer_signal_l, dd.peer_signal_r = left + land_bytes, right + land_bytes
    dd.peer_flags = flags_ptrs.data_ptr()
    dd.host_flag = hdev.value
    self._peer = dict(dd=dd, base=base, mapped=mapped, flags_ptrs=flags_ptrs, host=host.value)
    # the step kernels take their counts from the device ({N_LOC, N_ROWS} of `info`); the
    # descriptor they get carries the CAPACITY as n, so their grids never change and one
    # captured graph serves every rebuild
    nb = _lib.NbrT.from_buffer_copy(ws.c)
    nb.n = self.cap
    nb.n_rows = 0
    nb.n_dev = self.info.data_ptr() + 4 * _lib.DD_N_LOC
    self._nb_step = nb
    self._epoch = 0
    if W > 1:
      dist.barrier(group=self.comm.group)      # every block is mapped before the first push

  def close(self):
    """Unmaps the peers' blocks and frees this rank's (call on every rank)."""
    if self._peer is None:
      return
    torch.cuda.synchronize()
    if self.comm.world > 1:
      dist.barrier(group=self.comm.group)
    for rank, q in self._peer['mapped'].items():
      if rank != self.comm.rank:
        _lib.call('jmd_p2p_close', C.c_void_p(q))
    if self.comm.world > 1:
      dist.barrier(group=self.comm.group)
    _lib.call('jmd_p2p_free', C.c_void_p(self._peer['base']))
    _lib.call('jmd_host_flag_free', C.c_void_p(self._peer['host']))
    self._peer = None
    self._graph = None

  # -- pieces -----------------------------------------------------------------------
  def _select(self, st, thr_a, thr_b, list_a, list_b, counters):
    """Ascending int32 indices of owned atoms with d < thr_a / d >= thr_b into the
    fixed-capacity lists; counts stay on the device.  One ordered-select kernel
    (look-back scan): the atom order, hence the summation order, is reproducible."""
    _lib.call('jmd_dd_select_ordered', self.dtc, self.dim, self.cap, _lib.ptr(self.info), _lib.ptr(st.R),
              self.axis, float(self.lo), float(self.box[self.axis]), float(thr_a), float(thr_b),
              _lib.ptr(list_a), _lib.ptr(list_b), _lib.ptr(counters), list_a.numel(),
              _lib.ptr(self.sel_scratch), _lib.stream())

  def _migrate(self, st):
    """Atoms that left [lo, lo + width) move to the neighbouring rank (device
    side: select -> pack -> exchange -> compact; info[N_OWN] is updated)."""
    s = _lib.stream()
    self._select(st, 0.0, self.width, self.mig_a, self.mig_b, self.mig_counters)
    out_l, out_r, in_l, in_r = self.mig_pay
    g_out_l, g_out_r, g_in_l, g_in_r = self.mig_gid
    _lib.call('jmd_dd_pack_migrate', self.dtc, self.dim, self.cap_mig, _lib.ptr(self.mig_a),
              _lib.ptr(self.mig_b), _lib.ptr(self.mig_counters), _lib.ptr(st.R), _lib.ptr(st.P),
              _lib.ptr(st.F), _lib.ptr(st.gid), _lib.ptr(out_l), _lib.ptr(out_r), _lib.ptr(g_out_l),
              _lib.ptr(g_out_r), s)
    self.comm.exchange_many([(out_l, out_r, in_l, in_r), (g_out_l, g_out_r, g_in_l, g_in_r)])
    _lib.call('jmd_dd_compact', self.dtc, self.dim, self.cap, self.cap_mig, _lib.ptr(self.mig_a),
              _lib.ptr(self.mig_b), _lib.ptr(self.mig_counters), _lib.ptr(in_l), _lib.ptr(g_in_l),
              _lib.ptr(in_r), _lib.ptr(g_in_r), _lib.ptr(st.R), _lib.ptr(st.P), _lib.ptr(st.F),
              _lib.ptr(st.gid), _lib.ptr(self.mig_scratch), _lib.ptr(self.info), s)

  def _ghosts(self, st):
    """Select face atoms and exchange the ghost layer; ends with the one host
    read of a rebuild (`info`: owned / face / ghost counts and error bits)."""
    s = _lib.stream()
    self._select(st, self.ghost_width, self.width - self.ghost_width, self.list_a, self.list_b,
                 self.counters)
    out_l, out_r, in_l, in_r = self.face_msg
    for k, (idx, out) in enumerate(((self.list_a, out_l), (self.list_b, out_r))):
      _lib.call('jmd_dd_pack_counted', self.dtc, self.dim, self.cap_list, _lib.ptr(idx),
                _lib.ptr(self.counters[k:]), _lib.ptr(st.R), _lib.ptr(out), s)
    self.comm.exchange(out_l, out_r, in_l, in_r)
    _lib.call('jmd_dd_place', self.dtc, self.dim, self.cap, self.cap_list, _lib.ptr(self.counters),
              _lib.ptr(in_l), _lib.ptr(in_r), _lib.ptr(st.R), _lib.ptr(self.info), s)
    self.info_host.copy_(self.info, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    info = self.info_host.tolist()
    if info[_lib.DD_ERROR] & _lib.DD_ELIST:
      raise RuntimeError('domain decomposition list capacity exceeded')
    if info[_lib.DD_ERROR] & _lib.DD_ECAP:
      raise RuntimeError('slab capacity exceeded; raise capacity_factor')
    if info[_lib.DD_ERROR] & _lib.DD_ETIMEOUT:
      raise RuntimeError('a neighbour rank\'s halo or rebuild flag never arrived (peer exchange timed out)')
    st.n_own = info[_lib.DD_N_OWN]
    n_from_l, n_from_r = info[_lib.DD_FROM_L], info[_lib.DD_FROM_R]
    st.n_ghost = n_from_l + n_from_r
    self._lists = (self.list_a[:info[_lib.DD_FACE_L]], self.list_b[:info[_lib.DD_FACE_R]],
                   n_from_l, n_from_r)
    self.last_info = info

  def ghost_ids(self, st):
    """Global ids of the ghost atoms (diagnostics; one extra exchange)."""
    if self._lists is None:
      return st.gid[st.n_own:st.n_own]
    face_l, face_r, n_from_l, n_from_r = self._lists
    gl = st.gid[face_l.long()].contiguous()
    gr = st.gid[face_r.long()].contiguous()
    o = st.n_own
    self.comm.exchange(gl, gr, st.gid[o:o + n_from_l], st.gid[o + n_from_l:o + st.n_ghost])
    return st.gid[o:o + st.n_ghost]

  def _halo(self, st):
    """Face positions -> neighbours' ghost slots (every step; exact sizes known
    on the host since the last rebuild)."""
    if self._lists is None:
      return
    face_l, face_r, n_from_l, n_from_r = self._lists
    sl, sr = self._send_l[:face_l.numel()], self._send_r[:face_r.numel()]
    for idx, out in ((face_l, sl), (face_r, sr)):
      _lib.call('jmd_dd_pack', self.dtc, self.dim, idx.numel(), _lib.ptr(idx), _lib.ptr(st.R),
                _lib.ptr(out), _lib.stream())
    o = st.n_own
    self.comm.exchange(sl, sr, st.R[o:o + n_from_l], st.R[o + n_from_l:o + n_from_l + n_from_r])

  def _rebuild(self, st, first=False):
    if self.comm.world > 1:
      self._migrate(st)       # also fixes up atoms handed in slightly outside the slab
      self._ghosts(st)
    else:
      st.n_ghost = 0
      self._lists = None
    n_loc = st.n_own + st.n_ghost
    if self.comm.world == 1:
      # no ghost exchange ran: publish the counts the step kernels read
      self.info[_lib.DD_N_OWN] = st.n_own
      self.info[_lib.DD_N_LOC] = n_loc
      self.info[_lib.DD_N_ROWS] = st.n_own
      self.info[_lib.DD_FROM_L] = 0
      self.info[_lib.DD_FROM_R] = 0
      self.counters.zero_()
    Rl = st.R[:n_loc]
    if self.nbrs is None:
      self._allocate(st, Rl)
    else:
      ws = self.nbrs._ws
      ws.c.n, ws.c.n_rows = n_loc, st.n_own
      ws.n = n_loc
      s, pp = _lib.stream(), _lib.ptr(st.R)
      _lib.call('jmd_nbr_bin', ws.ref(), pp, 0, s)
      _lib.call('jmd_nbr_build', ws.ref(), pp, 0, 0, s)
      _lib.call('jmd_nbr_export', ws.ref(), pp, 0, s)
      # capacity overflow of the local list (density drifts through migration): caught at
      # the NEXT rebuild's host read (`_ghosts`), see _check_list
    if self._peer is not None or self.transport == 'p2p':
      # The graph-captured force kernel runs over the CAPACITY with n_rows = capacity: a
      # slot takes part iff perm[slot] < capacity.  Ghost slots have empty rows (cnt = 0,
      # set by the scan) and zero momenta, so they add nothing; slots behind the local
      # atoms must not alias a real atom.
      self.nbrs._ws.t['perm'][n_loc:].fill_(_INT_MAX)
      st.P[st.n_own:].zero_()
    self.rebuilds += 1

  def _allocate(self, st, Rl, extra=0):
    self.nbrs = self.neighbor_fn.allocate(Rl, extra_capacity=extra, n_capacity=self.cap,
                                          n_rows=st.n_own, no_public_idx=True)
    self._graph = None
    if self._peer is not None:
      ws = self.nbrs._ws
      nb = _lib.NbrT.from_buffer_copy(ws.c)
      nb.n, nb.n_rows = self.cap, 0
      nb.n_dev = self.info.data_ptr() + 4 * _lib.DD_N_LOC
      self._nb_step = nb
      self._peer['dd'].skin_blk = ws.t['skin_blk'].data_ptr()

  def check_list(self, st):
    """Host check of the local neighbour list's error bits (one sync): on overflow the
    list is re-allocated with more head-room from the current positions and the step
    graph is captured again.  Called from the bench / user loop every block of steps,
    like the reference's `did_buffer_overflow` check (partition.py:840-854)."""
    code = int(self.nbrs.error.code)
    if code & 3:
      n_loc = st.n_own + st.n_ghost
      self._extra = getattr(self, '_extra', 0) + 8
      self._allocate(st, st.R[:n_loc], extra=self._extra)
      self._force(st, kick=False)
      return True
    return False

  def _force(self, st, kick):
    ws = self.nbrs._ws
    fn = self.energy_fn
    _, species, params = fn._resolve(self.nbrs, {})
    ws.set_species(None)
    pt, keep, _ = fn._pair_struct(st.R, species, params, False)
    _lib.call('jmd_pair_force', ws.ref(), C.byref(pt), _lib.ptr(st.F), None,
              _lib.ptr(self.red), None, _lib.ptr(self.partials),
              _lib.ptr(st.P) if kick else None, _lib.ptr(self.mass), 0,
              self.dt_2, None, 0, _lib.stream())

  # -- the step ---------------------------------------------------------------------
  def _step_kernels(self, st):
    """drift -> push (halo + flag) -> wait (unpack + decision) -> force; device-side
    counts only, capturable."""
    s = _lib.stream()
    nb = C.byref(self._nb_step)
    dd = C.byref(self._peer['dd'])
    _lib.call('jmd_nve_kick_drift', C.byref(self.sp), self.dtc, self.cap, nb,
              _lib.ptr(st.R), _lib.ptr(st.P), _lib.ptr(st.F), _lib.ptr(self.mass), 0,
              self.dt, None, None, _lib.ptr(st.R), _lib.ptr(st.P), s)
    _lib.call('jmd_dd_comm_push', dd, _lib.ptr(st.R), s)
    _lib.call('jmd_dd_comm_wait', dd, nb, _lib.ptr(st.R), s)
    fn = self.energy_fn
    _, species, params = fn._resolve(self.nbrs, {})
    pt, keep, _ = fn._pair_struct(st.R, species, params, False)
    _lib.call('jmd_pair_force', nb, C.byref(pt), _lib.ptr(st.F), None,
              _lib.ptr(self.red), None, _lib.ptr(self.partials), _lib.ptr(st.P),
              _lib.ptr(self.mass), 0, self.dt_2, None, 0, s)

  def _poll(self, epoch, timeout=20.0):
    """Global rebuild decision of step `epoch` from the mapped host word (written by
    jmd_dd_comm_wait, i.e. BEFORE that step's force kernel runs)."""
    t0 = time.perf_counter()
    flag = self._host_flag
    while True:
      v = flag.value
      if (v >> 1) >= epoch:
        return bool(v & 1)
      if time.perf_counter() - t0 > timeout:
        raise RuntimeError(f'domain decomposition: no decision for step {epoch} '
                           f'(host word {v}); a peer rank stalled?')

  def _step_p2p(self, st):
    # NeighborList.update semantics (partition.py:1146) with a GLOBAL decision taken on
    # the positions the last step produced
    if self._pending is None and self._epoch > 0:
      self._pending = self._poll(self._epoch)
    if self._pending:
      self._rebuild(st)
    self._pending = None
    if self.use_graph and self.device.type == 'cuda' and self._warm >= 2:
      if self._graph is None:
        self._capture(st)
      self._graph.replay()
    else:
      self._step_kernels(st)      # the first steps run eagerly (lazy kernel loading
      self._warm += 1             # must not happen inside a capture)
    self._epoch += 1
    return st

  def _capture(self, st):
    """Captures the four step kernels once; every count they need is device-resident,
    so the graph survives rebuilds (it is dropped only when the list is re-allocated)."""
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      self._step_kernels(st)
    self._graph = g

  def step(self, st):
    if self._peer is not None:
      return self._step_p2p(st)
    return self._step_nccl(st)

  def _launch_decision(self, st):
    """Skin predicate on the CURRENT owned positions -> global OR -> pinned host
    flag (asynchronously).  Enqueued right after the drift, i.e. before the halo
    exchange and the force kernel of the same step, so by the time the host
    needs the answer (start of the next step) the GPU is still busy with the
    force kernel and never waits for the host."""
    ws = self.nbrs._ws
    main = torch.cuda.current_stream()
    side = self._side_stream if self.device.type == 'cuda' else None
    if not self._drift_flags:
      _lib.call('jmd_nbr_skin_check', ws.ref(), _lib.ptr(st.R), _lib.stream())
    # The reduction of the flags, their all-reduce across ranks and the copy to the
    # host run on a side stream: the main stream goes straight on to the halo
    # exchange and the force kernel instead of waiting for the collective.
    if side is not None:
      side.wait_stream(main)
    ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
    with ctx:
      if self._drift_flags:
        # the drift kernel just left one skin flag per 256 owned atoms (jmd_integrate.cu)
        nblk = (st.n_own + 255) // 256
        flag = ws.t['skin_blk'][:nblk].max().to(torch.int64).reshape(1)
        if ws.c.always_rebuild:
          flag = torch.ones_like(flag)
      else:
        flag = ws.t['state'][_lib.ST_REBUILD:_lib.ST_REBUILD + 1]
      if self.comm.world > 1:
        if self.comm.direct:
          self._flag_dev.copy_(flag)
          dist.all_reduce(self._flag_dev, op=dist.ReduceOp.MAX, group=self.comm.flag_group)
          self._flag_host.copy_(self._flag_dev, non_blocking=True)
        else:                                   # gloo: staged through the host
          self._flag_host.copy_(flag)
          dist.all_reduce(self._flag_host, op=dist.ReduceOp.MAX, group=self.comm.group)
      else:
        self._flag_host.copy_(flag, non_blocking=True)
      self._flag_event.record()
    self._decision_pending = True

  def _take_decision(self, st):
    if not self._decision_pending:
      self._drift_flags = False          # no drift produced these positions
      self._launch_decision(st)
    self._flag_event.synchronize()
    self._decision_pending = False
    if self.device.type == 'cuda':      # the side stream is idle now; order it before the next drift
      torch.cuda.current_stream().wait_stream(self._side_stream)
    return bool(int(self._flag_host[0]) != 0)

  def _step_nccl(self, st):
    """Host-orchestrated fallback (transport='nccl'; gloo staging for tests): one host
    decision per step, torch.distributed halo exchange."""
    s = _lib.stream()
    if self._take_decision(st):
      self._rebuild(st)
    ws = self.nbrs._ws
    _lib.call('jmd_nve_kick_drift', C.byref(self.sp), self.dtc, st.n_own, ws.ref(),
              _lib.ptr(st.R), _lib.ptr(st.P), _lib.ptr(st.F), _lib.ptr(self.mass), 0,
              self.dt, None, None, _lib.ptr(st.R), _lib.ptr(st.P), s)
    self._drift_flags = True
    self._launch_decision(st)
    if st.n_ghost:
      self._halo(st)
      _lib.call('jmd_nbr_pack_range', ws.ref(), _lib.ptr(st.R), st.n_own, st.n_ghost, s)
    self._force(st, kick=True)
    return st

  # -- observables ------------------------------------------------------------------
  def kinetic_energy(self):
    """Global KE of the last step (sum over ranks of the fused reduction)."""
    t = self.red[_lib.RED_KINETIC:_lib.RED_KINETIC + 1].clone()
    return float(self.comm.sum(t).item())

  def potential_energy(self, st):
    """Global potential energy at the current positions."""
    ws = self.nbrs._ws
    fn = self.energy_fn
    _, species, params = fn._resolve(self.nbrs, {})
    pt, keep, _ = fn._pair_struct(st.R, species, params, False)
    red = torch.zeros(_lib.RED_COUNT, dtype=torch.float64, device=self.device)
    Ftmp = torch.empty_like(st.F)
    _lib.call('jmd_nbr_pack', ws.ref(), _lib.ptr(st.R), _lib.stream())
    _lib.call('jmd_pair_force', ws.ref(), C.byref(pt), _lib.ptr(Ftmp), None, _lib.ptr(red),
              None, _lib.ptr(self.partials), None, None, 0, 0.0, None, 1, _lib.stream())
    t = red[_lib.RED_ENERGY:_lib.RED_ENERGY + 1].clone()
    return float(self.comm.sum(t).item())


# ----------------------------------------------------------------------------------
# bench.py entry for N > 1
# ----------------------------------------------------------------------------------

def untimed_steps(args):
  """Steps run before the timed window: the same count as the single-GPU arm (bench.py)."""
  import bench
  return bench.untimed_steps(args)


def measure_domain(args, comm, rank, world, dev, cells):
  """value / ms_per_step of `args.steps` steps for a slab of `cells` fcc cells per rank
  (stacked along x): the extra points of the multi-GPU JSON line."""
  import bench
  from . import energy
  R_loc, box_loc = bench.fcc(cells)
  a = box_loc[1] / cells[1]
  R_loc[:, 0] += rank * cells[0] * a
  box = np.array([world * cells[0] * a, cells[1] * a, cells[2] * a], np.float32)
  N_loc = len(R_loc)
  rng = np.random.default_rng(2000 + rank)
  P_loc = rng.normal(0, np.sqrt(bench.KT), (N_loc, 3)).astype(np.float32)
  disp, shift = space.periodic(box)
  _, efn = energy.lennard_jones_neighbor_list(disp, box, r_onset=2.0, r_cutoff=bench.R_CUT,
                                              dr_threshold=bench.SKIN)
  dom = SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, comm=comm)
  Rd, Pd = torch.as_tensor(R_loc, device=dev), torch.as_tensor(P_loc, device=dev)
  Pd -= (comm.sum(Pd.sum(0, dtype=torch.float64)) / (world * N_loc)).to(Pd.dtype)
  st = dom.init(Rd, Pd, torch.arange(N_loc, device=dev) + rank * N_loc)
  for _ in range(untimed_steps(args)):
    st = dom.step(st)
  torch.cuda.synchronize()
  dist.barrier()
  r0 = dom.rebuilds
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    st = dom.step(st)
  e1.record()
  torch.cuda.synchronize()
  dist.barrier()
  ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  n_tot = torch.tensor([st.n_own], dtype=torch.int64, device=dev)
  dist.all_reduce(n_tot)
  ov = torch.tensor([int(dom.nbrs.error.code) & 3], dtype=torch.int64, device=dev)
  dist.all_reduce(ov, op=dist.ReduceOp.MAX)
  out = {'atoms': int(n_tot.item()), 'atoms_per_gpu': N_loc,
         'value': int(n_tot.item()) * args.steps / (float(ms.item()) * 1e-3), 'unit': 'atom-timesteps/s',
         'ms_per_step': float(ms.item()) / args.steps, 'steps': args.steps,
         'rebuilds': dom.rebuilds - r0, 'neighbor_overflow': bool(int(ov.item())), 'n_gpus': world}
  dom.close()
  del dom, st
  torch.cuda.empty_cache()
  return out


def bench_domain(args, world, rank, dev):
  import time
  """Weak-scaling LJ NVE: every rank owns a slab of `cells^3` fcc cells stacked
  along x; prints the JSON line on rank 0 (bench.py contract)."""
  import json
  import bench
  from . import energy
  n = args.cells
  R_loc, box_loc = bench.fcc((n, n, n))
  a = box_loc[0] / n
  R_loc[:, 0] += rank * n * a
  box = np.array([world * n * a, n * a, n * a], np.float32)
  N_loc = len(R_loc)
  rng = np.random.default_rng(1000 + rank)
  P_loc = rng.normal(0, np.sqrt(bench.KT), (N_loc, 3)).astype(np.float32)
  comm = RingComm()
  disp, shift = space.periodic(box)
  _, efn = energy.lennard_jones_neighbor_list(disp, box, r_onset=2.0, r_cutoff=bench.R_CUT,
                                              dr_threshold=bench.SKIN)
  dom = SlabDomain(box, efn, bench.R_CUT, bench.SKIN, bench.DT, comm=comm)
  Rd = torch.as_tensor(R_loc, device=dev)
  Pd = torch.as_tensor(P_loc, device=dev)
  psum = comm.sum(Pd.sum(0, dtype=torch.float64))
  Pd -= (psum / (world * N_loc)).to(Pd.dtype)
  gid = torch.arange(N_loc, device=dev) + rank * N_loc
  st = dom.init(Rd, Pd, gid)
  for _ in range(untimed_steps(args)):
    st = dom.step(st)
  torch.cuda.synchronize()
  dist.barrier()
  r0 = dom.rebuilds
  sampler = bench.ClockSampler(dev.index or 0) if rank == 0 else None
  if sampler:
    sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize()
  dist.barrier()
  e0.record()
  for _ in range(args.steps):
    st = dom.step(st)
  e1.record()
  torch.cuda.synchronize()
  dist.barrier()
  ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  clocks = sampler.stop() if sampler else None
  n_tot = torch.tensor([st.n_own], dtype=torch.int64, device=dev)
  dist.all_reduce(n_tot)
  n_ghost = torch.tensor([st.n_ghost], dtype=torch.int64, device=dev)
  dist.all_reduce(n_ghost, op=dist.ReduceOp.MAX)
  ke = dom.kinetic_energy()
  rebuilds_timed = dom.rebuilds - r0
  dom_graph, dom_transport = bool(dom._graph is not None), ('p2p' if dom._peer is not None else 'nccl')

  # ---- dominant kernel (fused force + half kick) timed inside 20 more steps -------
  ws = dom.nbrs._ws
  pairs = int(torch.clamp(ws.t['cnt'][:st.n_own + st.n_ghost], max=ws.c.m_int).sum())
  kernel_bytes = pairs * 20 + st.n_own * 60
  evs = []
  plain_force = dom._force

  def timed_force(state, kick):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plain_force(state, kick=kick)
    b.record()
    evs.append((a, b))

  dom._force = timed_force
  graph_was, dom.use_graph = dom.use_graph, False        # (eager steps: the force launch is timed alone)
  if dom._peer is not None:
    real_kernels = dom._step_kernels

    def timed_kernels(state):
      # same four kernels, with events around the force launch
      s_ = _lib.stream()
      nb = C.byref(dom._nb_step)
      dd = C.byref(dom._peer['dd'])
      _lib.call('jmd_nve_kick_drift', C.byref(dom.sp), dom.dtc, dom.cap, nb, _lib.ptr(state.R), _lib.ptr(state.P),
                _lib.ptr(state.F), _lib.ptr(dom.mass), 0, dom.dt, None, None, _lib.ptr(state.R), _lib.ptr(state.P), s_)
      _lib.call('jmd_dd_comm_push', dd, _lib.ptr(state.R), s_)
      _lib.call('jmd_dd_comm_wait', dd, nb, _lib.ptr(state.R), s_)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      fn = dom.energy_fn
      _, species, params = fn._resolve(dom.nbrs, {})
      pt, keep, _ = fn._pair_struct(state.R, species, params, False)
      _lib.call('jmd_pair_force', nb, C.byref(pt), _lib.ptr(state.F), None, _lib.ptr(dom.red), None,
                _lib.ptr(dom.partials), _lib.ptr(state.P), _lib.ptr(dom.mass), 0, dom.dt_2, None, 0, s_)
      b.record()
      evs.append((a, b))
    dom._step_kernels = timed_kernels
  for _ in range(20):
    st = dom.step(st)
  torch.cuda.synchronize()
  dom._force = plain_force
  dom.use_graph = graph_was
  if dom._peer is not None:
    dom._step_kernels = real_kernels
  k_ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs[3:]]))],
                      dtype=torch.float64, device=dev)
  dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
  kb = torch.tensor([kernel_bytes], dtype=torch.int64, device=dev)
  dist.all_reduce(kb, op=dist.ReduceOp.MAX)

  # ---- end to end (same definition as the one-GPU line): the domain exists (allocation,
  # peer mapping and graph capture are the counterpart of the reference's first allocate +
  # jit compile); timed: H2D of the slab state from pinned memory, rebuild on it, first
  # force evaluation, the steps with the periodic host readbacks, D2H of the state.
  R_pin = torch.from_numpy(R_loc).pin_memory()
  P_pin = Pd.cpu().pin_memory()
  out_R = torch.empty_like(R_pin).pin_memory()
  out_P = torch.empty_like(P_pin).pin_memory()
  e2e_steps = args.steps
  torch.cuda.synchronize()
  dist.barrier()
  t0 = time.perf_counter()
  st = dom.reset(st, R_pin.to(dev, non_blocking=True), P_pin.to(dev, non_blocking=True), gid)
  for i in range(e2e_steps):
    st = dom.step(st)
    if (i + 1) % args.block == 0:
      dom.kinetic_energy()                       # global KE read back, like the example loop
      dom.check_list(st)
  n_out = min(st.n_own, N_loc)
  out_R[:n_out].copy_(st.R[:n_out], non_blocking=True)   # (slab populations drift by a few atoms)
  out_P[:n_out].copy_(st.P[:n_out], non_blocking=True)
  torch.cuda.synchronize()
  dist.barrier()
  e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
  dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
  overflow = torch.tensor([int(dom.nbrs.error.code) & 3], dtype=torch.int64, device=dev)
  dist.all_reduce(overflow, op=dist.ReduceOp.MAX)

  # ---- extra points: 4M atoms per GPU (N = 32M on 8 GPUs, the north star's run) and the
  # strong-scaling series (8M atoms in total, split over the ranks)
  extra = {}
  dom.close()
  del dom, st
  torch.cuda.empty_cache()
  if not args.no_extra:
    extra['weak_4m_per_gpu'] = measure_domain(args, comm, rank, world, dev, (100, 100, 100))
    if (8 * n) % world == 0:
      extra['strong_8m_total'] = measure_domain(args, comm, rank, world, dev, (8 * n // world, n, n))

  if rank == 0:
    ms_total = float(ms.item())
    N = int(n_tot.item())
    face_bytes = int(n_ghost.item()) * 12
    peak, peak_src = bench.peaks()
    kms = float(k_ms.item())
    achieved = float(kb.item()) / (kms * 1e-3) / 1e9
    line = {
        'metric': 'atom-timesteps/s', 'value': N * args.steps / (ms_total * 1e-3),
        'unit': 'atom-timesteps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'LJ fcc N={N} ({N_loc}/GPU) rho={bench.RHO} rc={bench.R_CUT} '
                               f'skin={bench.SKIN} dt={bench.DT} kT={bench.KT} NVE, slab decomposition along x',
                   'atoms': N, 'ghost_atoms_per_gpu': int(n_ghost.item()),
                   'halo_bytes_per_step_per_gpu': face_bytes,
                   'rebuilds_in_timed_region': rebuilds_timed,
                   'untimed_steps_before_window': untimed_steps(args),
                   'loop': ('CUDA graph of the 4 step kernels (drift, peer-memory push, wait+unpack, force); '
                            'host polls the device-written rebuild decision; rebuilds host-driven')
                   if dom_graph else 'eager loop',
                   'transport': dom_transport, 'neighbor_overflow': bool(int(overflow.item())),
                   'l2_policy': 'working set exceeds L2',
                   'kinetic_energy_per_atom': ke / N},
        'roofline': {'bound': 'hbm', 'kernel': 'k_pair_force<float,3,LJ,scalar,kick> (slowest rank)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': None, 'peak_source': peak_src, 'kernel_ms': kms,
                     'algorithmic_bytes_per_launch': int(kb.item())},
        'cpu_baseline': None,
        'e2e': {'value': N * e2e_steps / float(e2e_s.item()), 'unit': 'atom-timesteps/s',
                'h2d_bytes_per_step': 2 * N_loc * 12 * world / e2e_steps,
                'd2h_bytes_per_step': (2 * N_loc * 12 * world + 8 * world * (e2e_steps // args.block)) / e2e_steps,
                'steps': e2e_steps,
                'includes': 'per rank: H2D of the slab state from pinned memory, migration + ghost selection + '
                            f'neighbour build on it, first force evaluation, {e2e_steps} steps, global KE readback '
                            f'every {args.block} steps, D2H of state; domain allocation, peer mapping and graph '
                            'capture happen before the timed region (same definition as the one-GPU line)'},
        'weak_4m_per_gpu': extra.get('weak_4m_per_gpu'), 'strong_8m_total': extra.get('strong_8m_total'),
        'gpu_launches': int(world * (args.steps * 4 + rebuilds_timed * 22)),
        'clocks': clocks,
    }
    print(json.dumps(line))
  dist.destroy_process_group()
