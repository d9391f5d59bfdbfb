"""Enumeration factors (host mirror of pgmax/factor/enum.py:35-394).

The wiring keeps one *block* per factor group: all factors of a block share
one ``factor_configs`` table, so the table is stored once instead of being
expanded to the reference's ``factor_configs_edge_states[R, 2]`` (R = sum of
configs x arity; 1.07e9 rows for the 8192^2 Ising config).  The expanded array
is still produced on demand for the oracle and the wiring-equivalence tests.
"""

import dataclasses
import warnings
from typing import Any, Dict, Hashable, List, Mapping, Sequence, Tuple

import numpy as np

from pgmax_b200.factor import factor


@dataclasses.dataclass(frozen=True, eq=False)
class EnumBlock:
  """Factors sharing one configuration table, contiguous in the message vector.

  Attributes:
    num_factors: number of factors in the block.
    factor_configs: int array (num_configs, arity); entry = state of the a-th
      variable in that configuration.
    first_edge / first_config: offsets local to the enclosing EnumWiring
      (edge index of the block's first edge, config index of its first
      log-potential entry).
  """

  num_factors: int
  factor_configs: np.ndarray
  first_edge: int
  first_config: int

  @property
  def arity(self) -> int:
    return int(self.factor_configs.shape[1])

  @property
  def num_configs(self) -> int:
    return int(self.factor_configs.shape[0])


class EnumWiring(factor.Wiring):
  """Wiring of EnumFactors: per-edge table + blocks (+ lazy reference arrays)."""

  def __init__(self, edge_var_start, edge_num_states, edge_factor, blocks: Sequence[EnumBlock]):
    super().__init__(edge_var_start, edge_num_states, edge_factor)
    self.blocks = tuple(blocks)
    self._fces = None

  @property
  def num_val_configs(self) -> int:
    return sum(b.num_factors * b.num_configs for b in self.blocks)

  @property
  def factor_configs_edge_states(self) -> np.ndarray:
    """Reference layout [R, 2] = (config index, edge-state index), both local to
    the Enum slice; row-major over (factor, config, variable)
    (pgmax/factor/enum.py:364-394)."""
    if self._fces is None:
      msg_start = np.cumsum(self.edge_num_states) - self.edge_num_states
      parts = []
      for b in self.blocks:
        k, a = b.num_configs, b.arity
        starts = msg_start[b.first_edge : b.first_edge + b.num_factors * a]
        es = starts.reshape(b.num_factors, 1, a) + b.factor_configs[None]
        cfg = b.first_config + np.repeat(
            np.arange(b.num_factors * k, dtype=np.int64), a
        )
        parts.append(np.stack([cfg, es.reshape(-1)], axis=1))
      fces = (
          np.concatenate(parts, axis=0)
          if parts
          else np.empty((0, 2), dtype=np.int64)
      )
      fces.flags.writeable = False
      self._fces = fces
    return self._fces

  def get_inference_arguments(self) -> Dict[str, Any]:
    fces = self.factor_configs_edge_states
    return {
        "factor_configs_indices": fces[:, 0],
        "factor_configs_edge_states": fces[:, 1],
        "num_val_configs": self.num_val_configs,
        "num_factors": self.num_factors,
    }


@dataclasses.dataclass(frozen=True, eq=False)
class EnumFactor(factor.Factor):
  """Factor defined by an explicit list of valid configurations.

  Attributes:
    factor_configs: int array (num_val_configs, num_variables).
    log_potentials: float array (num_val_configs,).

  Validation and messages follow pgmax/factor/enum.py:95-133.
  """

  factor_configs: np.ndarray
  log_potentials: np.ndarray

  def __post_init__(self):
    self.factor_configs.flags.writeable = False
    if not np.issubdtype(self.factor_configs.dtype, np.integer):
      raise ValueError(
          f"Configurations should be integers. Got {self.factor_configs.dtype}."
      )
    if not np.issubdtype(self.log_potentials.dtype, np.floating):
      raise ValueError(
          f"Potential should be floats. Got {self.log_potentials.dtype}."
      )
    if self.factor_configs.ndim != 2:
      raise ValueError(
          "factor_configs should be a 2D array containing a list of valid"
          " configurations for EnumFactor. Got a factor_configs array of shape"
          f" {self.factor_configs.shape}."
      )
    if len(self.variables) != self.factor_configs.shape[1]:
      raise ValueError(
          f"Number of variables {len(self.variables)} doesn't match given"
          f" configurations {self.factor_configs.shape}"
      )
    if self.log_potentials.shape != (self.factor_configs.shape[0],):
      raise ValueError(
          "Expected log potentials of shape"
          f" {(self.factor_configs.shape[0],)} for"
          f" ({self.factor_configs.shape[0]}) valid configurations. Got log"
          f" potentials of shape {self.log_potentials.shape}."
      )
    limits = np.array([v[1] for v in self.variables])
    if ((self.factor_configs < 0) | (self.factor_configs >= limits[None])).any():
      raise ValueError("Invalid configurations for given variables")

  @staticmethod
  def concatenate_wirings(wirings: Sequence[EnumWiring]) -> EnumWiring:
    """Stacks EnumWirings (same role as pgmax/factor/enum.py:138-189)."""
    var_start, num_states, factors = factor.concatenate_edge_tables(wirings)
    blocks, edge_shift, config_shift = [], 0, 0
    for w in wirings:
      for b in w.blocks:
        blocks.append(
            dataclasses.replace(
                b,
                first_edge=b.first_edge + edge_shift,
                first_config=b.first_config + config_shift,
            )
        )
      edge_shift += w.num_edges
      config_shift += w.num_val_configs
    return EnumWiring(var_start, num_states, factors, blocks)

  @staticmethod
  def compile_wiring(
      factor_edges_num_states: np.ndarray,
      variables_for_factors: Sequence[List[Tuple[int, int]]],
      factor_configs: np.ndarray,
      vars_to_starts: Mapping[Tuple[int, int], int],
      num_factors: int,
  ) -> EnumWiring:
    """Wiring of one group of EnumFactors sharing ``factor_configs``.

    Same arguments and shape check as pgmax/factor/enum.py:192-270.
    """
    num_variables = factor_configs.shape[1]
    if factor_edges_num_states.shape != (num_factors * num_variables,):
      raise ValueError(
          "Expected factor_edges_num_states shape is"
          f" {(num_factors * num_variables,)}. Got"
          f" {factor_edges_num_states.shape}."
      )
    var_start, num_states, factors = factor.edge_table_for(
        variables_for_factors, vars_to_starts
    )
    # One block per run of consecutive factors whose variables have the same numbers of states:
    # inside a block the message / potential offsets are arithmetic progressions (pgx_enum_block).
    # The reference allows one group to span variables with different numbers of states
    # (tests/fgraph/test_fgraph.py:285-333); such a group becomes several blocks that share the
    # configuration table, in the reference's factor order.
    configs = np.ascontiguousarray(factor_configs, dtype=np.int64)
    signature = np.asarray(num_states, dtype=np.int64).reshape(num_factors, num_variables)
    breaks = np.flatnonzero(np.any(signature[1:] != signature[:-1], axis=1)) + 1
    starts = np.concatenate([[0], breaks]).astype(np.int64)
    ends = np.concatenate([breaks, [num_factors]]).astype(np.int64)
    blocks = [
        EnumBlock(
            num_factors=int(hi - lo),
            factor_configs=configs,
            first_edge=int(lo) * num_variables,
            first_config=int(lo) * int(configs.shape[0]),
        )
        for lo, hi in zip(starts, ends)
    ]
    return EnumWiring(var_start, num_states, factors, blocks)

  @staticmethod
  def compute_factor_energy(
      variables: List[Hashable],
      vars_to_map_states: Dict[Hashable, Any],
      factor_configs: np.ndarray,
      log_potentials: np.ndarray,
  ) -> float:
    """Energy of one EnumFactor under a decoding (pgmax/factor/enum.py:325-360)."""
    decoded = np.array([vars_to_map_states[v] for v in variables])
    hits = np.flatnonzero((np.asarray(factor_configs) == decoded[None]).all(axis=1))
    if hits.size == 0:
      warnings.warn(
          f"Invalid decoding for Enum factor {variables} "
          f"with variables set to {tuple(decoded.tolist())}!"
      )
      return float(np.inf)
    return float(-np.asarray(log_potentials)[hits[-1]])
