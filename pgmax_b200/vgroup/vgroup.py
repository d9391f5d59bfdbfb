"""Base class of variable groups (host mirror of pgmax/vgroup/vgroup.py:31).

A variable is the tuple ``(int64 hash, num_states)``.  The hash of the i-th
variable of a group is ``group_hash + i`` with ``group_hash`` a multiple of
MAX_SIZE, which is what lets the factor-graph builder turn a table of hashes
back into (group, index) pairs with integer arithmetic instead of dict
lookups (see fgraph.FactorGraph.var_starts_of).
"""

import itertools
from typing import Any, List, Tuple

import numpy as np

MAX_SIZE = 1e9
_GROUP_STRIDE = int(MAX_SIZE)
_group_counter = itertools.count(1)


class VarGroup:
  """Group of variables sharing one hash range.

  Attributes:
    num_states: int64 array, one entry per variable.
  """

  num_states: np.ndarray

  def _assign_hash(self) -> None:
    # Group ids are drawn from a process-wide counter: unique, ordered by
    # creation, and (unlike id()) never recycled.
    self._hash = next(_group_counter) * _GROUP_STRIDE
    assert self._hash < 2**63

  def __hash__(self) -> int:
    return self._hash

  def __eq__(self, other) -> bool:
    return isinstance(other, VarGroup) and self._hash == other._hash

  def __lt__(self, other) -> bool:
    return hash(self) < hash(other)

  def __getitem__(self, val: Any):
    raise NotImplementedError(
        "Please subclass the VarGroup class and override this method"
    )

  @property
  def variable_hashes(self) -> np.ndarray:
    raise NotImplementedError(
        "Please subclass the VarGroup class and override this method"
    )

  @property
  def variables(self) -> List[Tuple[int, int]]:
    """All variables of the group as (hash, num_states) tuples, C order."""
    cached = getattr(self, "_variables", None)
    if cached is None:
      cached = list(
          zip(
              self.variable_hashes.ravel().tolist(),
              self.num_states.ravel().tolist(),
          )
      )
      self._variables = cached
    return cached

  def flatten(self, data: Any) -> np.ndarray:
    raise NotImplementedError(
        "Please subclass the VarGroup class and override this method"
    )

  def unflatten(self, flat_data: np.ndarray, per_state: bool) -> Any:
    raise NotImplementedError(
        "Please subclass the VarGroup class and override this method"
    )
