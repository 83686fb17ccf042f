"""jax_md_b200: a B200-native (sm_100a CUDA) implementation of JAX MD's
short-range molecular-dynamics hot path behind the reference's own API
(`space`, `partition`, `smap`, `energy`, `quantity`, `simulate`, `minimize`).

Arrays are CUDA `torch.Tensor`s (PyTorch is plumbing: device memory, streams,
torch.distributed); all compute is hand-written CUDA in `libjmd_b200.so`,
reached through the C ABI in include/jmd_b200.h.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import dataclasses, util, space, partition, smap, energy, quantity  # noqa: F401
from . import simulate, minimize, lax, units  # noqa: F401

__version__ = '0.1.0'
