"""Factor graph and its compiled state (host mirror of pgmax/fgraph/fgraph.py:35-371)."""

import collections
import dataclasses
import types
from typing import Any, Dict, List, Mapping, OrderedDict, Sequence, Tuple, Type, Union

import numpy as np

from pgmax_b200 import factor
from pgmax_b200 import fgroup
from pgmax_b200 import vgroup
from pgmax_b200.factor import FACTOR_TYPES


@dataclasses.dataclass(frozen=True, eq=False)
class FactorGraphState:
  """Immutable compiled view of a FactorGraph (pgmax/fgraph/fgraph.py:35-78).

  Attributes:
    variable_groups: VarGroups in constructor order.
    vars_to_starts: variable -> first index in the flat evidence vector.
    num_var_states: length of the flat evidence vector (V_s).
    total_factor_num_states: length of the flat message vector (E_s).
    factor_type_to_msgs_range: factor type -> (start, end) in the messages.
    factor_type_to_potentials_range: factor type -> (start, end) in the potentials.
    factor_group_to_potentials_starts: factor group -> start in the potentials.
    factor_group_to_msgs_starts: factor group -> start in the messages.
    log_potentials: flat log potentials (C).
    evidence_to_vars: variable index of every evidence entry.
    wiring: factor type -> concatenated Wiring.
  """

  variable_groups: Sequence[vgroup.VarGroup]
  vars_to_starts: Mapping[Tuple[int, int], int]
  num_var_states: int
  total_factor_num_states: int
  factor_type_to_msgs_range: OrderedDict[Type[factor.Factor], Tuple[int, int]]
  factor_type_to_potentials_range: OrderedDict[Type[factor.Factor], Tuple[int, int]]
  factor_group_to_potentials_starts: OrderedDict[fgroup.FactorGroup, int]
  factor_group_to_msgs_starts: OrderedDict[fgroup.FactorGroup, int]
  log_potentials: np.ndarray
  evidence_to_vars: np.ndarray
  wiring: OrderedDict[Type[factor.Factor], factor.Wiring]

  def __post_init__(self):
    for field in dataclasses.fields(self):
      value = getattr(self, field.name)
      if isinstance(value, np.ndarray):
        value.flags.writeable = False
      elif isinstance(value, Mapping) and not isinstance(value, types.MappingProxyType):
        object.__setattr__(self, field.name, types.MappingProxyType(value))


class FactorGraph:
  """Variables + factor groups, bucketed by factor type.

  Args:
    variable_groups: one VarGroup or a list of VarGroups.
  """

  def __init__(self, variable_groups: Union[vgroup.VarGroup, Sequence[vgroup.VarGroup]]):
    if isinstance(variable_groups, vgroup.VarGroup):
      variable_groups = [variable_groups]
    self.variable_groups = list(variable_groups)
    self._factor_types_to_groups = collections.OrderedDict(
        (ft, []) for ft in FACTOR_TYPES
    )
    self._seen_variable_sets = {ft: set() for ft in FACTOR_TYPES}
    # Flat evidence layout: groups in order, variables in C order, num_states
    # consecutive slots each (pgmax/fgraph/fgraph.py:113-124).
    self._vars_to_starts: Dict[Tuple[int, int], int] = {}
    self._var_num_states = []
    offset = 0
    for vg in self.variable_groups:
      states = vg.num_states.reshape(-1)
      starts = offset + np.cumsum(states) - states
      self._vars_to_starts.update(zip(vg.variables, starts.tolist()))
      self._var_num_states.append(states)
      offset += int(states.sum())
    self._num_var_states = offset
    self._compiled = None

  def __hash__(self) -> int:
    return hash(
        tuple(g for groups in self._factor_types_to_groups.values() for g in groups)
    )

  def add_factors(self, factors) -> None:
    """Adds a Factor, a FactorGroup, or a list of them.

    Raises ValueError on duplicated variables inside a factor or on a second
    factor of the same type over the same variable set
    (pgmax/fgraph/fgraph.py:137-183).
    """
    if isinstance(factors, list):
      for item in factors:
        self.add_factors(item)
      return
    if isinstance(factors, fgroup.FactorGroup):
      group = factors
    elif isinstance(factors, factor.Factor):
      group = fgroup.SingleFactorGroup(
          variables_for_factors=[factors.variables], single_factor=factors
      )
    else:
      raise ValueError(f"Cannot add object of type {type(factors)} to a FactorGraph")
    factor_type = group.factor_type
    seen = self._seen_variable_sets[factor_type]
    for variables in group.variables_for_factors:
      key = frozenset(variables)
      if len(key) != len(variables):
        raise ValueError(
            f"A Factor of type {factor_type} involving variables"
            f" {variables} contains variables duplicates."
        )
      if key in seen:
        raise ValueError(
            f"A Factor of type {factor_type} involving variables"
            f" {key} already exists. Please merge the corresponding"
            " factors."
        )
      seen.add(key)
    self._factor_types_to_groups[factor_type].append(group)
    self._compiled = None

  @property
  def factor_groups(self) -> OrderedDict[Type[factor.Factor], List[fgroup.FactorGroup]]:
    return self._factor_types_to_groups

  @property
  def factors(self) -> OrderedDict[Type[factor.Factor], Tuple[factor.Factor, ...]]:
    return collections.OrderedDict(
        (ft, tuple(f for g in groups for f in g.factors))
        for ft, groups in self._factor_types_to_groups.items()
    )

  @property
  def evidence_to_vars(self) -> np.ndarray:
    states = (
        np.concatenate(self._var_num_states)
        if self._var_num_states
        else np.empty((0,), dtype=np.int64)
    )
    return np.repeat(np.arange(states.shape[0], dtype=np.int64), states)

  def _compile(self) -> FactorGraphState:
    msgs_range, pots_range = collections.OrderedDict(), collections.OrderedDict()
    group_msgs, group_pots = collections.OrderedDict(), collections.OrderedDict()
    wiring = collections.OrderedDict()
    potentials = []
    msg_cursor = pot_cursor = 0
    # Messages and potentials are ordered by factor type, then group, then
    # factor (pgmax/fgraph/fgraph.py:186-232).
    for ft, groups in self._factor_types_to_groups.items():
      msg_first, pot_first = msg_cursor, pot_cursor
      group_wirings = []
      for g in groups:
        group_msgs[g], group_pots[g] = msg_cursor, pot_cursor
        flat_lp = g.factor_group_log_potentials
        potentials.append(flat_lp)
        msg_cursor += int(g.factor_edges_num_states.sum())
        pot_cursor += int(flat_lp.shape[0])
        group_wirings.append(g.compile_wiring(self._vars_to_starts))
      msgs_range[ft] = (msg_first, msg_cursor)
      pots_range[ft] = (pot_first, pot_cursor)
      wiring[ft] = ft.concatenate_wirings(group_wirings)
    log_potentials = (
        np.concatenate(potentials).astype(np.float64)
        if potentials
        else np.empty((0,))
    )
    return FactorGraphState(
        variable_groups=self.variable_groups,
        vars_to_starts=self._vars_to_starts,
        num_var_states=self._num_var_states,
        total_factor_num_states=msg_cursor,
        factor_type_to_msgs_range=msgs_range,
        factor_type_to_potentials_range=pots_range,
        factor_group_to_potentials_starts=group_pots,
        factor_group_to_msgs_starts=group_msgs,
        log_potentials=log_potentials,
        evidence_to_vars=self.evidence_to_vars,
        wiring=wiring,
    )

  @property
  def fg_state(self) -> FactorGraphState:
    """Compiled state for the factors added so far (cached until add_factors)."""
    if self._compiled is None:
      self._compiled = self._compile()
    return self._compiled

  @property
  def wiring(self):
    return self.fg_state.wiring

  @property
  def log_potentials(self):
    state = self.fg_state
    return collections.OrderedDict(
        (ft, state.log_potentials[s:e])
        for ft, (s, e) in state.factor_type_to_potentials_range.items()
    )

  @property
  def bp_state(self) -> Any:
    """BPState with default potentials, zero messages and zero evidence."""
    from pgmax_b200.infer import bp_state as bps  # pylint: disable=g-import-not-at-top

    state = self.fg_state
    return bps.BPState(
        log_potentials=bps.LogPotentials(fg_state=state),
        ftov_msgs=bps.FToVMessages(fg_state=state),
        evidence=bps.Evidence(fg_state=state),
    )
