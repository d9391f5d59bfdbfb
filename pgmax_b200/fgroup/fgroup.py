"""Factor groups (host mirror of pgmax/fgroup/fgroup.py:30-296)."""

import collections
import inspect
from typing import Any, FrozenSet, List, Mapping, OrderedDict, Sequence, Tuple, Type

import numpy as np

from pgmax_b200.factor import factor as factor_lib


class FactorGroup:
  """A group of factors of one type.

  Attributes:
    variables_for_factors: one list of variables per factor.
    factor_configs: configuration table shared by the group (Enum only).
    log_potentials: array (num_factors, num_configs) (Enum only; empty otherwise).
    factor_type: the Factor subclass of every factor in the group.
  """

  factor_type: Type[Any]
  factor_configs = None

  def __init__(self, variables_for_factors: Sequence[List[Tuple[int, int]]]):
    if not variables_for_factors:
      raise ValueError("Cannot create a FactorGroup with no Factor.")
    self.variables_for_factors = variables_for_factors
    if not hasattr(self, "log_potentials"):
      self.log_potentials = np.empty((0,))
    self._cache = {}

  def __hash__(self):
    return id(self)

  def __eq__(self, other):
    return self is other

  def __lt__(self, other):
    return hash(self) < hash(other)

  def _memo(self, key, fn):
    if key not in self._cache:
      self._cache[key] = fn()
    return self._cache[key]

  def __getitem__(self, variables: Sequence[Tuple[int, int]]) -> Any:
    """The factor of the group connected to exactly this set of variables."""
    key = frozenset(variables)
    table = self._variables_to_factors
    if key not in table:
      raise ValueError(
          f"The queried factor connected to the set of variables {key} is"
          " not present in the factor group."
      )
    return table[key]

  @property
  def factor_sizes(self) -> np.ndarray:
    return self._memo(
        "factor_sizes",
        lambda: np.fromiter(
            (len(vs) for vs in self.variables_for_factors),
            dtype=np.int64,
            count=len(self.variables_for_factors),
        ),
    )

  @property
  def factor_edges_num_states(self) -> np.ndarray:
    """num_states of every (factor, variable) pair, factor-major."""
    return self._memo(
        "factor_edges_num_states",
        lambda: np.fromiter(
            (v[1] for vs in self.variables_for_factors for v in vs),
            dtype=np.int64,
            count=int(self.factor_sizes.sum()),
        ),
    )

  @property
  def _variables_to_factors(self) -> Mapping[FrozenSet[Any], factor_lib.Factor]:
    return self._memo("variables_to_factors", self._get_variables_to_factors)

  @property
  def factor_group_log_potentials(self) -> np.ndarray:
    """Flat log potentials of the group (factor-major)."""
    return self._memo(
        "flat_log_potentials", lambda: np.asarray(self.log_potentials).reshape(-1)
    )

  @property
  def factors(self) -> Tuple[factor_lib.Factor, ...]:
    return tuple(self._variables_to_factors.values())

  @property
  def num_factors(self) -> int:
    return len(self.variables_for_factors)

  def _get_variables_to_factors(self) -> OrderedDict[FrozenSet[Any], Any]:
    raise NotImplementedError(
        "Please subclass the FactorGroup class and override this method"
    )

  def flatten(self, data) -> np.ndarray:
    raise NotImplementedError(
        "Please subclass the FactorGroup class and override this method"
    )

  def unflatten(self, flat_data) -> Any:
    raise NotImplementedError(
        "Please subclass the FactorGroup class and override this method"
    )

  def compile_wiring(self, vars_to_starts: Mapping[Tuple[int, int], int]) -> Any:
    """Calls ``factor_type.compile_wiring`` with the group attributes it names
    (same introspection contract as pgmax/fgroup/fgroup.py:195-220)."""
    names = inspect.getfullargspec(self.factor_type.compile_wiring).args
    kwargs = {n: getattr(self, n) for n in names if n != "vars_to_starts"}
    return self.factor_type.compile_wiring(vars_to_starts=vars_to_starts, **kwargs)


class SingleFactorGroup(FactorGroup):
  """Wraps one Factor added directly to a FactorGraph (internal use)."""

  def __init__(self, variables_for_factors, single_factor: factor_lib.Factor):
    self.single_factor = single_factor
    self.factor_type = type(single_factor)
    lp = np.asarray(single_factor.log_potentials)
    # Group potentials are (num_factors, num_configs).
    self.log_potentials = lp[None] if lp.shape[0] > 0 else lp
    super().__init__(variables_for_factors)
    if len(self.variables_for_factors) != 1:
      raise ValueError(
          "SingleFactorGroup should only contain one factor. Got"
          f" {len(self.variables_for_factors)}"
      )
    names = inspect.getfullargspec(self.factor_type.compile_wiring).args
    for name in names:
      if name != "vars_to_starts" and not hasattr(self, name):
        setattr(self, name, getattr(single_factor, name))
    if hasattr(single_factor, "factor_configs"):
      self.factor_configs = single_factor.factor_configs

  def _get_variables_to_factors(self):
    return collections.OrderedDict(
        [(frozenset(self.variables_for_factors[0]), self.single_factor)]
    )

  def flatten(self, data):
    raise NotImplementedError(
        "SingleFactorGroup does not support vectorized factor operations."
    )

  def unflatten(self, flat_data):
    raise NotImplementedError(
        "SingleFactorGroup does not support vectorized factor operations."
    )
