"""N-dimensional grid of variables (host mirror of pgmax/vgroup/varray.py:28)."""

from typing import List, Tuple, Union

import numpy as np

from pgmax_b200.vgroup import vgroup


def _as_host(data):
  """Brings array-likes (numpy, torch, lists) to a numpy array without copying when possible."""
  if hasattr(data, "detach"):  # torch tensor
    return data.detach().cpu().numpy()
  return np.asarray(data)


class NDVarArray(vgroup.VarGroup):
  """Variables laid out on a grid of the given shape.

  Args:
    num_states: int (shared) or int array of shape ``shape`` (ragged).
    shape: grid shape.

  Same constructor contract and error messages as the reference
  (pgmax/vgroup/varray.py:40-70).
  """

  def __init__(self, num_states: Union[int, np.ndarray], shape: Tuple[int, ...]):
    self.shape = tuple(int(s) for s in shape)
    self._assign_hash()
    max_size = int(vgroup.MAX_SIZE)
    if np.prod(self.shape) > max_size:
      raise ValueError(
          f"Currently only support NDVarArray of size smaller than {max_size}."
          f" Got {np.prod(self.shape)}"
      )
    if np.isscalar(num_states):
      self.num_states = np.full(self.shape, num_states, dtype=np.int64)
    elif isinstance(num_states, np.ndarray) and np.issubdtype(
        num_states.dtype, np.integer
    ):
      if num_states.shape != self.shape:
        raise ValueError(
            f"Expected num_states shape {self.shape}. Got {num_states.shape}."
        )
      self.num_states = num_states.astype(np.int64)
    else:
      raise ValueError(
          "num_states should be an integer or a NumPy array of dtype int"
      )
    self.num_states.flags.writeable = False
    self._size = int(np.prod(self.shape))
    self._uniform = (
        self._size == 0
        or int(self.num_states.min()) == int(self.num_states.max())
    )

  def __repr__(self) -> str:
    if self.num_states.size == 0:
      desc = f"num_states={self.num_states}"
    else:
      lo, hi = int(self.num_states.min()), int(self.num_states.max())
      desc = (
          f"num_states={lo}"
          if lo == hi
          else f"min_num_states={lo}, max_num_states={hi}"
      )
    return f"NDVarArray(shape={self.shape}, {desc}, _hash={self._hash})"

  @property
  def variable_hashes(self) -> np.ndarray:
    cached = getattr(self, "_hashes", None)
    if cached is None:
      cached = self._hash + np.arange(self._size, dtype=np.int64).reshape(
          self.shape
      )
      cached.flags.writeable = False
      self._hashes = cached
    return cached

  def __getitem__(
      self, val
  ) -> Union[Tuple[int, int], List[Tuple[int, int]]]:
    """vg[i, j] -> one variable; vg[slices] -> list of variables (C order).

    Out-of-range indices raise IndexError through numpy, as in the reference
    (pgmax/vgroup/varray.py:92-117).
    """
    names = self.variable_hashes[val]
    states = self.num_states[val]
    if isinstance(names, np.ndarray):
      return list(zip(names.ravel().tolist(), states.ravel().tolist()))
    return (int(names), int(states))

  # ---------------------------------------------------------------- flatten
  def _state_mask(self) -> np.ndarray:
    smax = int(self.num_states.max(initial=0))
    return np.arange(smax) < self.num_states[..., None]

  def flatten(self, data) -> np.ndarray:
    """Structured -> flat.  Accepts ``shape`` (per variable) or ``shape + (max_states,)``
    (per state); one extra leading axis is treated as a batch axis (the
    replacement for jax.vmap over bp.init)."""
    data = _as_host(data)
    smax = int(self.num_states.max(initial=0))
    per_var, per_state = self.shape, self.shape + (smax,)
    if data.shape == per_var:
      return data.reshape(-1)
    if data.shape == per_state:
      return data.reshape(-1) if self._uniform else data[self._state_mask()]
    if data.ndim >= 1 and data.shape[1:] == per_var:
      return data.reshape(data.shape[0], -1)
    if data.ndim >= 1 and data.shape[1:] == per_state:
      if self._uniform:
        return data.reshape(data.shape[0], -1)
      return data[:, self._state_mask()]
    raise ValueError(
        f"data should be of shape {per_var} or {per_state}. Got {data.shape}."
    )

  def unflatten(self, flat_data, per_state: bool) -> np.ndarray:
    """Flat -> structured; ragged groups are padded with -inf (reference
    pgmax/vgroup/varray.py:186-198).  A 2-D input is treated as (batch, flat)."""
    flat_data = _as_host(flat_data)
    if flat_data.ndim == 2:
      return np.stack([self.unflatten(row, per_state) for row in flat_data])
    if flat_data.ndim != 1:
      raise ValueError(
          f"Can only unflatten 1D array. Got a {flat_data.ndim}D array."
      )
    if per_state:
      total = int(self.num_states.sum())
      if flat_data.size != total:
        raise ValueError(
            f"flat_data size should be equal to {total}. Got size"
            f" {flat_data.size}."
        )
      smax = int(self.num_states.max(initial=0))
      if self._uniform:
        return flat_data.reshape(self.shape + (smax,))
      out = np.full(self.shape + (smax,), -np.inf, dtype=flat_data.dtype)
      out[self._state_mask()] = flat_data
      return out
    if flat_data.size != self._size:
      raise ValueError(
          f"flat_data size should be equal to {self._size}. "
          f"Got size {flat_data.size}."
      )
    return flat_data.reshape(self.shape)
