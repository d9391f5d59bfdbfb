"""Groups of OR / AND factors (host mirror of pgmax/fgroup/logical.py:28-95)."""

import collections

from pgmax_b200.factor import logical
from pgmax_b200.fgroup import fgroup


class LogicalFactorGroup(fgroup.FactorGroup):
  """Factors of one logical subtype; ``edge_states_offset`` is +1 (OR) or -1 (AND)."""

  edge_states_offset: int

  def _get_variables_to_factors(self):
    return collections.OrderedDict(
        (frozenset(vs), self.factor_type(variables=vs))
        for vs in self.variables_for_factors
    )


class ORFactorGroup(LogicalFactorGroup):
  edge_states_offset = 1
  factor_type = logical.ORFactor


class ANDFactorGroup(LogicalFactorGroup):
  edge_states_offset = -1
  factor_type = logical.ANDFactor
