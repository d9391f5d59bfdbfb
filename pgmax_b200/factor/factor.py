"""Factor / Wiring base classes (host mirror of pgmax/factor/factor.py:29-225).

Unlike the reference, a Wiring stores the incidence *per edge* (one row per
(factor, variable) pair) rather than per edge-state: an edge's states are
contiguous both in the message vector and in the evidence vector, so
``(first var-state, num_states, factor)`` per edge carries the same
information as the reference's ``var_states_for_edges[E_s, 3]`` at a fraction
of the size.  The reference-format array is still available (lazily expanded)
because the oracle and the wiring-equivalence tests use it.
"""

import dataclasses
from typing import Any, Dict, Hashable, List, Mapping, Sequence, Tuple

import numpy as np


def expand_edge_table(
    edge_var_start: np.ndarray,
    edge_num_states: np.ndarray,
    edge_factor: np.ndarray,
) -> np.ndarray:
  """Per-edge table -> reference layout ``var_states_for_edges[E_s, 3]``.

  Column 0: global var-state, column 1: edge index, column 2: factor index
  (pgmax/factor/factor.py:29-39, filled there by the numba loop at :197-225).
  """
  num_edges = edge_var_start.shape[0]
  out = np.empty((int(edge_num_states.sum()), 3), dtype=np.int64)
  if num_edges == 0:
    return out
  edge_ids = np.repeat(np.arange(num_edges, dtype=np.int64), edge_num_states)
  msg_starts = np.cumsum(edge_num_states) - edge_num_states
  state_in_edge = np.arange(out.shape[0], dtype=np.int64) - msg_starts[edge_ids]
  out[:, 0] = edge_var_start[edge_ids] + state_in_edge
  out[:, 1] = edge_ids
  out[:, 2] = edge_factor[edge_ids]
  return out


class Wiring:
  """Incidence structure of all factors of one type.

  Attributes:
    edge_var_start: int64[num_edges], var-state index of state 0 of each edge.
    edge_num_states: int64[num_edges].
    edge_factor: int64[num_edges], factor index local to this wiring.
    var_states_for_edges: int64[E_s, 3] in the reference layout (lazy).
  """

  def __init__(self, edge_var_start, edge_num_states, edge_factor):
    self.edge_var_start = np.ascontiguousarray(edge_var_start, dtype=np.int64)
    self.edge_num_states = np.ascontiguousarray(edge_num_states, dtype=np.int64)
    self.edge_factor = np.ascontiguousarray(edge_factor, dtype=np.int64)
    for arr in (self.edge_var_start, self.edge_num_states, self.edge_factor):
      arr.flags.writeable = False
    self._vsfe = None

  @property
  def num_edges(self) -> int:
    return int(self.edge_var_start.shape[0])

  @property
  def num_edge_states(self) -> int:
    return int(self.edge_num_states.sum())

  @property
  def num_factors(self) -> int:
    return int(self.edge_factor[-1]) + 1 if self.num_edges else 0

  @property
  def var_states_for_edges(self) -> np.ndarray:
    if self._vsfe is None:
      vsfe = expand_edge_table(
          self.edge_var_start, self.edge_num_states, self.edge_factor
      )
      vsfe.flags.writeable = False
      self._vsfe = vsfe
    return self._vsfe

  def get_inference_arguments(self) -> Dict[str, Any]:
    raise NotImplementedError(
        "Please subclass the Wiring class and override this method."
    )


def concatenate_edge_tables(
    wirings: Sequence[Wiring],
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
  """Stacks per-edge tables, shifting factor ids (edge ids are positional)."""
  if not wirings:
    empty = np.empty((0,), dtype=np.int64)
    return empty, empty, empty
  factor_shift = 0
  factors = []
  for w in wirings:
    factors.append(w.edge_factor + factor_shift)
    factor_shift += w.num_factors
  return (
      np.concatenate([w.edge_var_start for w in wirings]),
      np.concatenate([w.edge_num_states for w in wirings]),
      np.concatenate(factors),
  )


def concatenate_var_states_for_edges(
    list_var_states_for_edges: Sequence[np.ndarray],
) -> np.ndarray:
  """Reference-layout concatenation with edge / factor offsets.

  Same behaviour as pgmax/factor/factor.py:132-192: None inputs raise
  ValueError, empty blocks are skipped, columns 1 and 2 are shifted by the
  running number of edges / factors.
  """
  if list_var_states_for_edges is None:
    raise ValueError("list_var_states_for_edges cannot be None")
  blocks, edge_shift, factor_shift = [], 0, 0
  for block in list_var_states_for_edges:
    if block is None:
      raise ValueError("var_states_for_edges cannot be None")
    if block.shape[0] == 0:
      continue
    blocks.append(block + np.array([[0, edge_shift, factor_shift]], dtype=np.int64))
    edge_shift += int(block[-1, 1]) + 1
    factor_shift += int(block[-1, 2]) + 1
  if not blocks:
    return np.empty((0, 3), dtype=np.int64)
  return np.concatenate(blocks, axis=0)


def edge_table_for(
    variables_for_factors: Sequence[List[Tuple[int, int]]],
    vars_to_starts: Mapping[Tuple[int, int], int],
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
  """(first var-state, num_states, factor id) for every (factor, variable) pair."""
  starts = [vars_to_starts[v] for vs in variables_for_factors for v in vs]
  states = [v[1] for vs in variables_for_factors for v in vs]
  sizes = np.fromiter(
      (len(vs) for vs in variables_for_factors),
      dtype=np.int64,
      count=len(variables_for_factors),
  )
  factor_ids = np.repeat(np.arange(sizes.shape[0], dtype=np.int64), sizes)
  return (
      np.asarray(starts, dtype=np.int64),
      np.asarray(states, dtype=np.int64),
      factor_ids,
  )


@dataclasses.dataclass(frozen=True, eq=False)
class Factor:
  """A single factor.

  Attributes:
    variables: variables connected by the factor, each ``(hash, num_states)``.
    log_potentials: log potentials (empty for logical / pool factors).
  """

  variables: List[Tuple[int, int]]
  log_potentials: np.ndarray

  def __post_init__(self):
    if not hasattr(self, "compile_wiring"):
      raise NotImplementedError(
          "Please implement compile_wiring in for your factor"
      )

  @staticmethod
  def concatenate_wirings(wirings: Sequence[Wiring]) -> Wiring:
    raise NotImplementedError(
        "Please subclass the Wiring class and override this method."
    )

  @staticmethod
  def compute_factor_energy(
      variables: List[Hashable], vars_to_map_states: Dict[Hashable, Any]
  ) -> float:
    raise NotImplementedError(
        "Please subclass the Factor class and override this method"
    )
