"""OR / AND factors (host mirror of pgmax/factor/logical.py:39-488)."""

import dataclasses
import warnings
from typing import Any, Dict, Hashable, List, Mapping, Sequence, Tuple

import numpy as np

from pgmax_b200.factor import factor

# Below this temperature the sum-product update adds max-based lower bounds
# (pgmax/factor/logical.py:33).
TEMPERATURE_STABILITY_THRE = 0.5


class LogicalWiring(factor.Wiring):
  """Wiring of ORFactors or ANDFactors.

  Attributes:
    parents_edge_states: int64[P, 2] = (factor index, message index of the
      parent's state 0 for OR / state 1 for AND); the parent's other state is
      at ``+ edge_states_offset``.
    children_edge_states: int64[F], same convention for the child variable.
    edge_states_offset: +1 for OR, -1 for AND.
  Indices are local to the OR (or AND) slice of the message vector
  (pgmax/factor/logical.py:39-74).
  """

  def __init__(
      self,
      edge_var_start,
      edge_num_states,
      edge_factor,
      parents_edge_states,
      children_edge_states,
      edge_states_offset: int,
  ):
    super().__init__(edge_var_start, edge_num_states, edge_factor)
    self.parents_edge_states = np.ascontiguousarray(
        parents_edge_states, dtype=np.int64
    ).reshape(-1, 2)
    self.children_edge_states = np.ascontiguousarray(
        children_edge_states, dtype=np.int64
    )
    self.parents_edge_states.flags.writeable = False
    self.children_edge_states.flags.writeable = False
    self.edge_states_offset = int(edge_states_offset)

  def _validate(self) -> None:
    """Checks of pgmax/factor/logical.py:78-97."""
    num_factors = self.children_edge_states.shape[0]
    if num_factors == 0:
      return
    factor_ids = self.parents_edge_states[:, 0]
    if np.unique(factor_ids).shape[0] != num_factors:
      raise ValueError(
          f"The LogicalWiring must have {num_factors} different"
          " LogicalFactor indices"
      )
    if factor_ids.max() >= num_factors:
      raise ValueError(
          f"The highest LogicalFactor index must be {num_factors - 1}"
      )
    if self.edge_states_offset not in (1, -1):
      raise ValueError(
          "The LogicalWiring's edge_states_offset must be 1 (for OR) and -1"
          f" (for AND), but is {self.edge_states_offset}"
      )

  def get_inference_arguments(self) -> Dict[str, Any]:
    self._validate()
    return {
        "parents_factor_indices": self.parents_edge_states[:, 0],
        "parents_msg_indices": self.parents_edge_states[:, 1],
        "children_edge_states": self.children_edge_states,
        "edge_states_offset": self.edge_states_offset,
    }


def _parents_and_children(
    factor_sizes: np.ndarray, relevant_state: int
) -> Tuple[np.ndarray, np.ndarray]:
  """Message indices of parents / children for binary factors laid out back to back.

  Factor f with n_f variables owns messages [2*sum_{g<f} n_g, +2 n_f); its last
  variable is the child (pgmax/factor/logical.py:458-488).
  """
  factor_sizes = np.asarray(factor_sizes, dtype=np.int64)
  num_parents = factor_sizes - 1
  factor_msg_start = 2 * (np.cumsum(factor_sizes) - factor_sizes)
  factor_ids = np.repeat(np.arange(factor_sizes.shape[0], dtype=np.int64), num_parents)
  parent_rank = np.arange(factor_ids.shape[0], dtype=np.int64) - np.repeat(
      np.cumsum(num_parents) - num_parents, num_parents
  )
  parents = np.stack(
      [factor_ids, factor_msg_start[factor_ids] + 2 * parent_rank + relevant_state],
      axis=1,
  )
  children = factor_msg_start + 2 * num_parents + relevant_state
  return parents, children


@dataclasses.dataclass(frozen=True, eq=False)
class LogicalFactor(factor.Factor):
  """OR/AND factor (p1, ..., pn, c): parents first, child last; all binary."""

  log_potentials: np.ndarray = dataclasses.field(
      init=False, default_factory=lambda: np.empty((0,))
  )
  edge_states_offset: int = dataclasses.field(init=False)

  def __post_init__(self):
    if len(self.variables) < 2:
      raise ValueError(
          "A LogicalFactor requires at least one parent variable and one child"
          " variable"
      )
    if any(v[1] != 2 for v in self.variables):
      raise ValueError("All the variables in a LogicalFactor should be binary")

  @staticmethod
  def concatenate_wirings(wirings: Sequence[LogicalWiring]) -> LogicalWiring:
    """Stacks LogicalWirings (same role as pgmax/factor/logical.py:140-206)."""
    if not wirings:
      empty = np.empty((0,), dtype=np.int64)
      return LogicalWiring(
          empty, empty, empty, np.empty((0, 2), dtype=np.int64), empty, 1
      )
    var_start, num_states, factors = factor.concatenate_edge_tables(wirings)
    parents, children, factor_shift, msg_shift = [], [], 0, 0
    for w in wirings:
      parents.append(w.parents_edge_states + np.array([[factor_shift, msg_shift]]))
      children.append(w.children_edge_states + msg_shift)
      factor_shift += w.children_edge_states.shape[0]
      msg_shift += w.num_edge_states
    return LogicalWiring(
        var_start,
        num_states,
        factors,
        np.concatenate(parents, axis=0),
        np.concatenate(children, axis=0),
        wirings[0].edge_states_offset,
    )

  @staticmethod
  def compile_wiring(
      factor_edges_num_states: np.ndarray,
      variables_for_factors: Sequence[List[Tuple[int, int]]],
      factor_sizes: np.ndarray,
      vars_to_starts: Mapping[Tuple[int, int], int],
      edge_states_offset: int,
  ) -> LogicalWiring:
    """Wiring of a group of OR (offset +1) or AND (offset -1) factors.

    Same arguments as pgmax/factor/logical.py:209-291.
    """
    del factor_edges_num_states  # all variables are binary
    var_start, num_states, factors = factor.edge_table_for(
        variables_for_factors, vars_to_starts
    )
    relevant_state = (1 - edge_states_offset) // 2
    parents, children = _parents_and_children(factor_sizes, relevant_state)
    return LogicalWiring(
        var_start, num_states, factors, parents, children, edge_states_offset
    )


@dataclasses.dataclass(frozen=True, eq=False)
class ORFactor(LogicalFactor):
  """F(p1..pn, c) = 0 iff c == OR(p1..pn), -inf otherwise."""

  edge_states_offset: int = dataclasses.field(init=False, default=1)

  @staticmethod
  def compute_factor_energy(
      variables: List[Hashable], vars_to_map_states: Dict[Hashable, Any], **kwargs
  ) -> float:
    del kwargs  # factor_configs / log_potentials: not used (as in the reference)
    states = np.array([vars_to_map_states[v] for v in variables])
    if bool(np.any(states[:-1])) != bool(states[-1]):
      warnings.warn(
          f"Invalid decoding for OR factor {variables} "
          f"with parents set to {states[:-1]} "
          f"and child set to {states[-1]}!"
      )
      return float(np.inf)
    return 0.0


@dataclasses.dataclass(frozen=True, eq=False)
class ANDFactor(LogicalFactor):
  """F(p1..pn, c) = 0 iff c == AND(p1..pn), -inf otherwise."""

  edge_states_offset: int = dataclasses.field(init=False, default=-1)

  @staticmethod
  def compute_factor_energy(
      variables: List[Hashable], vars_to_map_states: Dict[Hashable, Any], **kwargs
  ) -> float:
    del kwargs  # factor_configs / log_potentials: not used (as in the reference)
    states = np.array([vars_to_map_states[v] for v in variables])
    if bool(np.all(states[:-1])) != bool(states[-1]):
      warnings.warn(
          f"Invalid decoding for AND factor {variables} "
          f"with parents set to {states[:-1]} "
          f"and child set to {states[-1]}!"
      )
      return float(np.inf)
    return 0.0
