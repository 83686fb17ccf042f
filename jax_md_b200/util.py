"""dtype helpers (reference jax_md/util.py)."""
import numpy as np
import torch

f32 = np.float32
f64 = np.float64
i32 = np.int32


def is_array(x):
  return isinstance(x, (np.ndarray, torch.Tensor))


def to_numpy(x):
  if isinstance(x, torch.Tensor):
    return x.detach().cpu().numpy()
  return x


def maybe_downcast(x):
  """util.py:109-112: keep f64 arrays, everything else becomes f32."""
  if isinstance(x, torch.Tensor):
    return x if x.dtype == torch.float64 else x.to(torch.float32)
  if isinstance(x, np.ndarray) and x.dtype == np.float64:
    return x
  if isinstance(x, np.generic) and x.dtype == np.float64:
    return x
  return np.asarray(x, np.float32)[()]


def np_max(x):
  return np.max(to_numpy(x))


def torch_dtype(np_or_torch_dtype):
  if isinstance(np_or_torch_dtype, torch.dtype):
    return np_or_torch_dtype
  return {np.dtype('float32'): torch.float32,
          np.dtype('float64'): torch.float64}[np.dtype(np_or_torch_dtype)]


def high_precision_sum(X, axis=None, keepdims=False):
  """util.py:91-106."""
  return X.sum(dim=axis, keepdim=keepdims, dtype=torch.float64).to(X.dtype) \
      if axis is not None else X.sum(dtype=torch.float64).to(X.dtype)
