"""`fori_loop`: the host-side counterpart of `jit(lax.fori_loop(...))`, which is
how JAX MD users run the hot path (examples/nve_neighbor_list.py:186-195):

    def body(i, carry):
      state, nbrs = carry
      nbrs = nbrs.update(state.position)
      state = apply_fn(state, neighbor=nbrs)
      return state, nbrs
    state, nbrs = lax.fori_loop(0, 100, body, (state, nbrs))

XLA turns that loop into one device program; here the same body is captured once
into a CUDA graph (`unroll` iterations per graph) and replayed, so the ~75 us of
Python/ctypes dispatch per step disappears -- which is what bounds small systems
(N <= 1e5).  Nothing in `update` / `apply_fn` synchronises with the host, so they
are capture-safe; the rebuild decision stays on the device.

The carry may be any nesting of tuples / lists / dicts / this package's
dataclasses; tensor leaves are carried through static buffers, everything else
(neighbour workspaces, callables, Python scalars) is treated as static.
"""
import dataclasses as _dc

import torch

from . import _lib


def _flatten(obj, leaves, path=()):
  """Returns a structure with tensor leaves replaced by their index."""
  if isinstance(obj, torch.Tensor):
    leaves.append(obj)
    return ('leaf', len(leaves) - 1)
  if isinstance(obj, tuple):
    return ('tuple', [_flatten(o, leaves) for o in obj])
  if isinstance(obj, list):
    return ('list', [_flatten(o, leaves) for o in obj])
  if isinstance(obj, dict):
    return ('dict', {k: _flatten(v, leaves) for k, v in obj.items()})
  if _dc.is_dataclass(obj) and not isinstance(obj, type) and not hasattr(obj, '_ws'):
    # (a NeighborList -- `_ws` -- is updated in place and carried by identity)
    fields = {}
    for f in _dc.fields(obj):
      if f.metadata.get('static', False):
        continue
      fields[f.name] = _flatten(getattr(obj, f.name), leaves)
    return ('dataclass', obj, fields)
  return ('static', obj)        # NeighborList (in-place workspace), scalars, fns


def _inplace_buffers(obj, out):
  """Tensors that kernels update in place behind objects carried by identity
  (e.g. the Nose-Hoover chain buffer)."""
  if isinstance(obj, (tuple, list)):
    for o in obj:
      _inplace_buffers(o, out)
  elif isinstance(obj, dict):
    for o in obj.values():
      _inplace_buffers(o, out)
  elif _dc.is_dataclass(obj) and not isinstance(obj, type):
    buf = getattr(obj, '_buf', None)
    if isinstance(buf, torch.Tensor):
      out.append(buf)
    elif not hasattr(obj, '_ws'):
      for f in _dc.fields(obj):
        _inplace_buffers(getattr(obj, f.name), out)


def _rebuild(struct, leaves):
  kind = struct[0]
  if kind == 'leaf':
    return leaves[struct[1]]
  if kind == 'tuple':
    return tuple(_rebuild(s, leaves) for s in struct[1])
  if kind == 'list':
    return [_rebuild(s, leaves) for s in struct[1]]
  if kind == 'dict':
    return {k: _rebuild(v, leaves) for k, v in struct[1].items()}
  if kind == 'dataclass':
    _, obj, fields = struct
    return _dc.replace(obj, **{k: _rebuild(v, leaves) for k, v in fields.items()})
  return struct[1]


class GraphLoop:
  """`unroll` iterations of `body_fun` captured in one CUDA graph that feeds its
  outputs back into its own static inputs; `run(k)` replays it k times."""

  def __init__(self, body_fun, init_val, unroll=10, warmup=2):
    _lib.require_cuda()
    self.unroll = unroll
    leaves = []
    self.struct = _flatten(init_val, leaves)
    self.static = [t.clone() for t in leaves]
    inplace = []
    _inplace_buffers(init_val, inplace)
    saved = [t.clone() for t in inplace]
    # eager warm-up on a side stream (lazy initialisations must not be captured)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      val = _rebuild(self.struct, [t.clone() for t in leaves])
      for i in range(warmup):
        val = body_fun(i, val)
    torch.cuda.current_stream().wait_stream(s)
    for t, keep in zip(inplace, saved):      # undo the warm-up on in-place state
      t.copy_(keep)
    torch.cuda.synchronize()
    # the warm-up advanced in-place state (neighbour workspaces); the caller's
    # init_val tensors are untouched and are what the graph starts from
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      val = _rebuild(self.struct, self.static)
      for i in range(unroll):
        val = body_fun(i, val)
      out = []
      out_struct = _flatten(val, out)
      if len(out) != len(self.static):
        raise ValueError('body_fun must return a carry with the same structure')
      for dst, src in zip(self.static, out):
        if dst.shape != src.shape or dst.dtype != src.dtype:
          raise ValueError('body_fun changed the shape/dtype of a carried tensor')
        if dst.data_ptr() != src.data_ptr():
          dst.copy_(src)
    self.out_struct = out_struct

  def load(self, val):
    leaves = []
    _flatten(val, leaves)
    for dst, src in zip(self.static, leaves):
      dst.copy_(src)

  def run(self, replays):
    for _ in range(replays):
      self.graph.replay()

  def value(self, clone=True):
    leaves = [t.clone() for t in self.static] if clone else self.static
    return _rebuild(self.out_struct, leaves)


def fori_loop(lower, upper, body_fun, init_val, unroll=10, graph=None):
  """`jax.lax.fori_loop(lower, upper, body_fun, init_val)` on a CUDA graph.

  Returns the final carry.  Pass `graph=` (a `GraphLoop` from a previous call,
  available as `fori_loop.last`) to reuse the captured graph across calls.  The
  loop index seen by `body_fun` during capture is the index within the graph,
  not the global one: bodies must not depend on it."""
  n = int(upper) - int(lower)
  if n <= 0:
    return init_val
  if n < unroll or not torch.cuda.is_available():
    val = init_val
    for i in range(lower, upper):
      val = body_fun(i, val)
    return val
  g = graph
  if g is None:
    # capturing warms up in place (neighbour lists are rebuilt on the warm-up
    # positions); loading init_val afterwards restores the carried tensors, and
    # the first replayed update() rebuilds the list if it has to.
    g = GraphLoop(body_fun, init_val, unroll=unroll)
  g.load(init_val)
  g.run(n // unroll)
  val = g.value()
  for i in range(n % unroll):
    val = body_fun(i, val)
  fori_loop.last = g
  return val


fori_loop.last = None
