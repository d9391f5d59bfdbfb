"""Groups of EnumFactors (host mirror of pgmax/fgroup/enum.py:30-435)."""

import collections
from typing import Optional

import numpy as np

from pgmax_b200.factor import enum as enum_lib
from pgmax_b200.fgroup import fgroup
from pgmax_b200.vgroup.varray import _as_host


def _reshape_flat(flat_data, options):
  """Reshapes a 1-D array to the first shape in ``options`` whose size matches."""
  flat_data = _as_host(flat_data)
  if flat_data.ndim != 1:
    raise ValueError(
        f"Can only unflatten 1D array. Got a {flat_data.ndim}D array."
    )
  for shape in options:
    if flat_data.size == int(np.prod(shape)):
      return flat_data.reshape(shape)
  raise ValueError(
      "flat_data should be compatible with shape"
      f" {options[0]} or {options[1]}. Got {flat_data.shape}."
  )


class EnumFactorGroup(fgroup.FactorGroup):
  """EnumFactors sharing one set of valid configurations.

  Args:
    variables_for_factors: one list of variables per factor.
    factor_configs: int array (num_val_configs, num_variables).
    log_potentials: None (zeros), (num_val_configs,) (shared) or
      (num_factors, num_val_configs).
  """

  factor_type = enum_lib.EnumFactor

  def __init__(self, variables_for_factors, factor_configs: np.ndarray, log_potentials: Optional[np.ndarray] = None):
    super().__init__(variables_for_factors)
    self.factor_configs = factor_configs
    num_configs = factor_configs.shape[0]
    full = (self.num_factors, num_configs)
    if log_potentials is None:
      log_potentials = np.zeros(full, dtype=float)
    else:
      log_potentials = _as_host(log_potentials)
      if log_potentials.shape not in ((num_configs,), full):
        raise ValueError(
            f"Expected log potentials shape: {(num_configs,)} or"
            f" {full}. Got {log_potentials.shape}."
        )
      log_potentials = np.broadcast_to(log_potentials, full)
    if not np.issubdtype(log_potentials.dtype, np.floating):
      raise ValueError(
          f"Potentials should be floats. Got {log_potentials.dtype}."
      )
    self.log_potentials = log_potentials

  def _get_variables_to_factors(self):
    lp = np.array(self.log_potentials)
    return collections.OrderedDict(
        (
            frozenset(vs),
            enum_lib.EnumFactor(
                variables=vs, factor_configs=self.factor_configs, log_potentials=lp[i]
            ),
        )
        for i, vs in enumerate(self.variables_for_factors)
    )

  def _message_width(self) -> int:
    return int(sum(v[1] for v in self.variables_for_factors[0]))

  def flatten(self, data) -> np.ndarray:
    """(num_configs,) | (F, num_configs) | (F, num_edge_states) -> flat
    (pgmax/fgroup/enum.py:118-156); one extra leading axis = batch.  The reference's shapes
    are matched FIRST: a (B, num_configs) array with B == num_factors is per-factor data, not a
    batch of shared rows (pass (B, F, num_configs) to say "batch" unambiguously)."""
    data = _as_host(data)
    nf, nc, width = self.num_factors, self.factor_configs.shape[0], self._message_width()
    if data.shape == (nc,):
      return np.tile(data, nf)
    if data.shape in ((nf, nc), (nf, width)):
      return data.reshape(-1)
    if data.ndim == 2 and data.shape[1:] == (nc,):
      return np.tile(data, (1, nf))
    if data.ndim == 3 and data.shape[1:] in ((nf, nc), (nf, width)):
      return data.reshape(data.shape[0], -1)
    raise ValueError(
        f"data should be of shape {(nf, nc)} or {(nf, width)} or {(nc,)}. Got"
        f" {data.shape}."
    )

  def unflatten(self, flat_data) -> np.ndarray:
    nf = self.num_factors
    return _reshape_flat(
        flat_data,
        [(nf, self.factor_configs.shape[0]), (nf, self._message_width())],
    )


class PairwiseFactorGroup(fgroup.FactorGroup):
  """EnumFactors over two variables where every joint state is valid.

  Args:
    variables_for_factors: one [var0, var1] list per factor.
    log_potential_matrix: None (zeros), (n0, n1) (shared) or (num_factors, n0, n1).
  """

  factor_type = enum_lib.EnumFactor

  def __init__(self, variables_for_factors, log_potential_matrix: Optional[np.ndarray] = None):
    super().__init__(variables_for_factors)
    first = self.variables_for_factors[0]
    if log_potential_matrix is None:
      log_potential_matrix = np.zeros((first[0][1], first[1][1]))
    else:
      log_potential_matrix = _as_host(log_potential_matrix)
    if log_potential_matrix.ndim not in (2, 3):
      raise ValueError(
          "log_potential_matrix should be either a 2D array, specifying shared"
          " parameters for all pairwise factors, or 3D array, specifying"
          " parameters for individual pairwise factors. Got a"
          f" {log_potential_matrix.ndim}D log_potential_matrix array."
      )
    if not np.issubdtype(log_potential_matrix.dtype, np.floating):
      raise ValueError(
          f"Potential matrix should be floats. Got {log_potential_matrix.dtype}."
      )
    if (
        log_potential_matrix.ndim == 3
        and log_potential_matrix.shape[0] != self.num_factors
    ):
      raise ValueError(
          f"Expected log_potential_matrix for {self.num_factors} factors. Got"
          f" log_potential_matrix for {log_potential_matrix.shape[0]} factors."
      )
    pair_shape = tuple(log_potential_matrix.shape[-2:])
    sizes = self.factor_sizes
    if (sizes != 2).any():
      bad = self.variables_for_factors[int(np.flatnonzero(sizes != 2)[0])]
      raise ValueError(
          "All pairwise factors should connect to exactly 2 variables. Got a"
          f" factor connecting to {len(bad)} variables ({bad})."
      )
    states = self.factor_edges_num_states.reshape(-1, 2)
    mismatch = np.flatnonzero((states != np.array(pair_shape)[None]).any(axis=1))
    if mismatch.size:
      bad = self.variables_for_factors[int(mismatch[0])]
      raise ValueError(
          f"The specified pairwise factor {bad} (with"
          f" {(bad[0][1], bad[1][1])}configurations) does not match the specified"
          f" log_potential_matrix (with {pair_shape} configurations)."
      )
    self.log_potential_matrix = log_potential_matrix
    n0, n1 = pair_shape
    # Row-major enumeration of all (state0, state1) pairs.
    self.factor_configs = np.stack(
        [np.repeat(np.arange(n0), n1), np.tile(np.arange(n1), n0)], axis=1
    )
    self.log_potentials = np.broadcast_to(
        log_potential_matrix, (self.num_factors,) + pair_shape
    ).reshape(self.num_factors, n0 * n1)

  def _get_variables_to_factors(self):
    return collections.OrderedDict(
        (
            frozenset(vs),
            enum_lib.EnumFactor(
                variables=vs,
                factor_configs=self.factor_configs,
                log_potentials=np.array(self.log_potentials[i]),
            ),
        )
        for i, vs in enumerate(self.variables_for_factors)
    )

  def flatten(self, data) -> np.ndarray:
    """(n0, n1) | (F, n0, n1) | (F, n0 + n1) -> flat (pgmax/fgroup/enum.py:345-381);
    one extra leading axis = batch.  The reference's shapes are matched first: a (B, n0, n1)
    array with B == num_factors is per-factor data (pass (B, F, n0, n1) for a batch)."""
    data = _as_host(data)
    nf = self.num_factors
    pair = tuple(self.log_potential_matrix.shape[-2:])
    width = (nf, int(sum(pair)))
    if data.shape == pair:
      return np.tile(data.reshape(-1), nf)
    if data.shape in ((nf,) + pair, width):
      return data.reshape(-1)
    if data.ndim >= 3 and data.shape[1:] == pair:
      return np.tile(data.reshape(data.shape[0], -1), (1, nf))
    if data.ndim >= 3 and data.shape[1:] in ((nf,) + pair, width):
      return data.reshape(data.shape[0], -1)
    raise ValueError(
        f"data should be of shape {(nf,) + pair} or {width} or {pair}. Got"
        f" {data.shape}."
    )

  def unflatten(self, flat_data) -> np.ndarray:
    nf = self.num_factors
    pair = tuple(self.log_potential_matrix.shape[-2:])
    return _reshape_flat(flat_data, [(nf,) + pair, (nf, int(sum(pair)))])
