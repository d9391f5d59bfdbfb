"""Pool factors (host mirror of pgmax/factor/pool.py:34-239)."""

import dataclasses
import warnings
from typing import Any, Dict, Hashable, List, Mapping, Sequence, Tuple

import numpy as np

from pgmax_b200.factor import factor
from pgmax_b200.factor import logical


class PoolWiring(factor.Wiring):
  """Wiring of PoolFactors.

  Attributes:
    pool_choices_edge_states: int64[P, 2] = (factor index, message index of the
      pool choice's state 0); state 1 is at +1.
    pool_indicators_edge_states: int64[F], message index of the indicator's
      state 0.
  (pgmax/factor/pool.py:34-60.)
  """

  def __init__(
      self,
      edge_var_start,
      edge_num_states,
      edge_factor,
      pool_choices_edge_states,
      pool_indicators_edge_states,
  ):
    super().__init__(edge_var_start, edge_num_states, edge_factor)
    self.pool_choices_edge_states = np.ascontiguousarray(
        pool_choices_edge_states, dtype=np.int64
    ).reshape(-1, 2)
    self.pool_indicators_edge_states = np.ascontiguousarray(
        pool_indicators_edge_states, dtype=np.int64
    )
    self.pool_choices_edge_states.flags.writeable = False
    self.pool_indicators_edge_states.flags.writeable = False

  def get_inference_arguments(self) -> Dict[str, Any]:
    """Checks of pgmax/factor/pool.py:64-77, then the kernel arguments."""
    num_factors = self.pool_indicators_edge_states.shape[0]
    if self.pool_choices_edge_states.shape[0] > 0:
      factor_ids = self.pool_choices_edge_states[:, 0]
      if np.unique(factor_ids).shape[0] != num_factors:
        raise ValueError(
            f"The PoolWiring must have {num_factors} different"
            " PoolFactor indices"
        )
      if factor_ids.max() >= num_factors:
        raise ValueError(
            f"The highest PoolFactor index must be {num_factors - 1}"
        )
    return {
        "pool_choices_factor_indices": self.pool_choices_edge_states[:, 0],
        "pool_choices_msg_indices": self.pool_choices_edge_states[:, 1],
        "pool_indicators_edge_states": self.pool_indicators_edge_states,
    }


@dataclasses.dataclass(frozen=True, eq=False)
class PoolFactor(factor.Factor):
  """F(pc1..pcn, pi) = 0 iff all zero, or pi = 1 and exactly one pc is 1."""

  log_potentials: np.ndarray = dataclasses.field(
      init=False, default_factory=lambda: np.empty((0,))
  )

  def __post_init__(self):
    if len(self.variables) < 2:
      raise ValueError(
          "A PoolFactor requires at least one pool choice and one pool "
          "indicator."
      )
    if any(v[1] != 2 for v in self.variables):
      raise ValueError("All the variables in a PoolFactor should all be binary")

  @staticmethod
  def concatenate_wirings(wirings: Sequence[PoolWiring]) -> PoolWiring:
    """Stacks PoolWirings (same role as pgmax/factor/pool.py:117-149)."""
    as_logical = [
        logical.LogicalWiring(
            w.edge_var_start,
            w.edge_num_states,
            w.edge_factor,
            w.pool_choices_edge_states,
            w.pool_indicators_edge_states,
            1,
        )
        for w in wirings
    ]
    lw = logical.LogicalFactor.concatenate_wirings(as_logical)
    return PoolWiring(
        lw.edge_var_start,
        lw.edge_num_states,
        lw.edge_factor,
        lw.parents_edge_states,
        lw.children_edge_states,
    )

  @staticmethod
  def compile_wiring(
      factor_edges_num_states: np.ndarray,
      variables_for_factors: Sequence[List[Tuple[int, int]]],
      factor_sizes: np.ndarray,
      vars_to_starts: Mapping[Tuple[int, int], int],
  ) -> PoolWiring:
    """Wiring of a group of PoolFactors; reuses the logical layout with offset +1
    (pgmax/factor/pool.py:153-180)."""
    lw = logical.LogicalFactor.compile_wiring(
        factor_edges_num_states=factor_edges_num_states,
        variables_for_factors=variables_for_factors,
        factor_sizes=factor_sizes,
        vars_to_starts=vars_to_starts,
        edge_states_offset=1,
    )
    return PoolWiring(
        lw.edge_var_start,
        lw.edge_num_states,
        lw.edge_factor,
        lw.parents_edge_states,
        lw.children_edge_states,
    )

  @staticmethod
  def compute_factor_energy(
      variables: List[Hashable], vars_to_map_states: Dict[Hashable, Any], **kwargs
  ) -> float:
    del kwargs  # factor_configs / log_potentials: not used (as in the reference)
    states = np.array([vars_to_map_states[v] for v in variables])
    if int(np.sum(states[:-1])) != int(states[-1]):
      warnings.warn(
          f"Invalid decoding for Pool factor {variables} "
          f"with pool choices set to {states[:-1]} "
          f"and pool indicator set to {states[-1]}!"
      )
      return float(np.inf)
    return 0.0
