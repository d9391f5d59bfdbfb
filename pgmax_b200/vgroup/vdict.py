"""Named set of variables (host mirror of pgmax/vgroup/vdict.py:27)."""

from typing import Any, Dict, Hashable, Mapping, Tuple, Union

import numpy as np

from pgmax_b200.vgroup import vgroup
from pgmax_b200.vgroup.varray import _as_host


class VarDict(vgroup.VarGroup):
  """Variables addressed by arbitrary hashable names.

  Args:
    num_states: int (shared) or int array of shape (len(variable_names),).
    variable_names: tuple of names.
  """

  def __init__(self, num_states: Union[int, np.ndarray], variable_names: Tuple[Any, ...]):
    self.variable_names = tuple(variable_names)
    self._assign_hash()
    n = len(self.variable_names)
    if np.isscalar(num_states):
      self.num_states = np.full((n,), num_states, dtype=np.int64)
    elif isinstance(num_states, np.ndarray) and np.issubdtype(
        num_states.dtype, np.integer
    ):
      if num_states.shape != (n,):
        raise ValueError(
            f"Expected num_states shape ({n},). Got {num_states.shape}."
        )
      self.num_states = num_states.astype(np.int64)
    else:
      raise ValueError(
          "num_states should be an integer or a NumPy array of dtype int"
      )
    self.num_states.flags.writeable = False
    self._name_to_idx = {name: i for i, name in enumerate(self.variable_names)}

  @property
  def variable_hashes(self) -> np.ndarray:
    return self._hash + np.arange(len(self.variable_names), dtype=np.int64)

  def __getitem__(self, var_name: Any) -> Tuple[int, int]:
    idx = self._name_to_idx.get(var_name)
    if idx is None:
      raise ValueError(f"Variable {var_name} is not in VarDict")
    return (self._hash + idx, int(self.num_states[idx]))

  def flatten(self, data: Mapping[Any, Any]) -> np.ndarray:
    """Dict name -> array of shape (num_states,) or (1,)  ->  flat vector
    (reference pgmax/vgroup/vdict.py:88-127)."""
    for var_name in data:
      idx = self._name_to_idx.get(var_name)
      if idx is None:
        raise ValueError(
            f"data is referring to a non-existent variable {var_name}."
        )
      shape = np.shape(data[var_name])
      if shape != (int(self.num_states[idx]),) and shape != (1,):
        raise ValueError(
            f"Variable {var_name} expects a data array of shape"
            f" {(int(self.num_states[idx]),)} or (1,). Got {shape}."
        )
    return np.concatenate(
        [_as_host(data[name]).reshape(-1) for name in self.variable_names]
    )

  def unflatten(
      self, flat_data, per_state: bool
  ) -> Dict[Hashable, np.ndarray]:
    """Flat vector -> dict (reference pgmax/vgroup/vdict.py:129-189; 1-D input only, as
    the reference - the batch axis goes through unflatten_batch)."""
    flat_data = _as_host(flat_data)
    if flat_data.ndim != 1:
      raise ValueError(
          f"Can only unflatten 1D array. Got a {flat_data.ndim}D array."
      )
    return self._unflatten(flat_data, per_state)

  def unflatten_batch(self, flat_data, per_state: bool) -> Dict[Hashable, np.ndarray]:
    """[B, flat] -> dict of arrays with a leading batch axis (the batch axis that replaces
    jax.vmap; used by InfererContext.get_beliefs / unflatten_states)."""
    flat_data = _as_host(flat_data)
    if flat_data.ndim != 2:
      raise ValueError(f"Can only unflatten a (batch, flat) array. Got a {flat_data.ndim}D array.")
    return self._unflatten(flat_data, per_state)

  def _unflatten(self, flat_data, per_state: bool) -> Dict[Hashable, np.ndarray]:
    num_variables = len(self.variable_names)
    if per_state:
      total = int(self.num_states.sum())
      if flat_data.shape[-1] != total:
        raise ValueError(
            "flat_data should be shape "
            f"(num_variable_states(={total}),). Got "
            f"{flat_data.shape}"
        )
      bounds = np.concatenate([[0], np.cumsum(self.num_states)])
      return {
          name: flat_data[..., bounds[i] : bounds[i + 1]]
          for i, name in enumerate(self.variable_names)
      }
    if flat_data.shape[-1] != num_variables:
      raise ValueError(
          "flat_data should be shape "
          f"(num_variables(={num_variables}),). Got "
          f"{flat_data.shape}"
      )
    return {name: flat_data[..., i] for i, name in enumerate(self.variable_names)}
