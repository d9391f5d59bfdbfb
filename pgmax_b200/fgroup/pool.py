"""Groups of Pool factors (host mirror of pgmax/fgroup/pool.py:28-61)."""

import collections

from pgmax_b200.factor import pool
from pgmax_b200.fgroup import fgroup


class PoolFactorGroup(fgroup.FactorGroup):
  factor_type = pool.PoolFactor

  def _get_variables_to_factors(self):
    return collections.OrderedDict(
        (frozenset(vs), pool.PoolFactor(variables=vs))
        for vs in self.variables_for_factors
    )
