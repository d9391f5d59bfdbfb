"""BP state containers (host mirror of pgmax/infer/bp_state.py:30-436).

``BPArrays`` is the data contract of the hot path: three flat fp32 vectors
(log potentials [C], factor->variable messages [E_s], evidence [V_s]).  A
leading batch axis on any of them replaces ``jax.vmap`` over ``bp.init`` /
``bp.run``; arrays without it are shared by the whole batch.
Fields may be numpy arrays (host) or torch CUDA tensors (device-resident).
"""

import dataclasses
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

from pgmax_b200 import fgraph
from pgmax_b200 import fgroup
from pgmax_b200.vgroup.varray import _as_host


def _is_torch(x) -> bool:
  return hasattr(x, "detach") and hasattr(x, "device")


@dataclasses.dataclass(frozen=True, eq=False)
class BPArrays:
  """The three flat arrays of belief propagation.

  Attributes:
    log_potentials: [C] or [B, C].
    ftov_msgs: [E_s] or [B, E_s].
    evidence: [V_s] or [B, V_s].
  """

  log_potentials: Any
  ftov_msgs: Any
  evidence: Any

  def __post_init__(self):
    for field in dataclasses.fields(self):
      value = getattr(self, field.name)
      if isinstance(value, np.ndarray) and value.flags.writeable:
        # a read-only VIEW: the caller's own buffer stays writable (the reference's arrays are
        # immutable jnp arrays, pgmax/infer/bp_state.py:45-48; ours alias user memory)
        view = value.view()
        view.flags.writeable = False
        object.__setattr__(self, field.name, view)

  @property
  def batch_size(self) -> Optional[int]:
    """Leading batch size, or None when all three arrays are 1-D."""
    sizes = {
        int(a.shape[0])
        for a in (self.log_potentials, self.ftov_msgs, self.evidence)
        if a.ndim == 2
    }
    if len(sizes) > 1:
      raise ValueError(f"Inconsistent batch sizes in BPArrays: {sorted(sizes)}")
    return sizes.pop() if sizes else None


def _set_slice(flat: np.ndarray, start: int, data: np.ndarray) -> np.ndarray:
  """flat[..., start:start+n] = data, growing a batch axis on demand."""
  data = np.asarray(data, dtype=np.float32)
  if data.ndim == 2 and flat.ndim == 1:
    flat = np.broadcast_to(flat, (data.shape[0],) + flat.shape).copy()
  elif data.ndim == 2 and flat.shape[0] != data.shape[0]:
    raise ValueError(
        f"Batch size mismatch: array has {flat.shape[0]}, update has {data.shape[0]}."
    )
  flat[..., start : start + data.shape[-1]] = data
  return flat


def _writable(arr) -> np.ndarray:
  return np.array(_as_host(arr), dtype=np.float32)


def update_log_potentials(
    log_potentials, updates: Dict[Any, Any], fg_state: fgraph.FactorGraphState
) -> np.ndarray:
  """New log potentials with per-FactorGroup updates (pgmax/infer/bp_state.py:58-101)."""
  out = _writable(log_potentials)
  for group, data in updates.items():
    if group not in fg_state.factor_group_to_potentials_starts:
      raise ValueError("Invalid FactorGroup for log potentials updates.")
    flat = group.flatten(data)
    expected = group.factor_group_log_potentials.shape
    if flat.shape[-1:] != expected:
      raise ValueError(
          f"Expected log potentials shape {expected} for"
          f" factor group. Got incompatible data shape {np.shape(data)}."
      )
    out = _set_slice(out, fg_state.factor_group_to_potentials_starts[group], flat)
  return out


def update_ftov_msgs(
    ftov_msgs, updates: Dict[Any, Any], fg_state: fgraph.FactorGraphState
) -> np.ndarray:
  """New messages with updates keyed by factor type (whole slice) or by variable
  (value spread evenly over the variable's edges) (pgmax/infer/bp_state.py:172-234)."""
  out = _writable(ftov_msgs)
  edge_starts = None
  for name, data in updates.items():
    data = _as_host(data)
    if name in fg_state.factor_type_to_msgs_range:
      start, end = fg_state.factor_type_to_msgs_range[name]
      if data.shape[-1:] != (end - start,) or data.ndim > 2:
        raise ValueError(
            f"Expected ftov_msgs shape {(end - start,)}"
            f" for factor type {name}. Got incompatible shape {data.shape}."
        )
      out = _set_slice(out, start, data)
    elif name in fg_state.vars_to_starts:
      if data.shape[-1:] != (name[1],) or data.ndim > 2:
        raise ValueError(
            f"Expected ftov_msgs shape {(name[1],)} for variable {name}."
            f" Got incompatible shape {data.shape}."
        )
      if edge_starts is None:
        # Global message start and first var-state of every edge, all types.
        var_start = np.concatenate(
            [w.edge_var_start for w in fg_state.wiring.values()]
        )
        num_states = np.concatenate(
            [w.edge_num_states for w in fg_state.wiring.values()]
        )
        edge_starts = (var_start, np.cumsum(num_states) - num_states)
      hits = edge_starts[1][edge_starts[0] == fg_state.vars_to_starts[name]]
      for start in hits.tolist():
        out = _set_slice(out, start, data / hits.shape[0])
    else:
      raise ValueError(
          "Provided variable or factor type is not in the FactorGraph"
      )
  return out


def update_evidence(
    evidence, updates: Dict[Any, Any], fg_state: fgraph.FactorGraphState
) -> np.ndarray:
  """New evidence with updates keyed by VarGroup or by variable
  (pgmax/infer/bp_state.py:292-342)."""
  out = _writable(evidence)
  groups = {id(g): g for g in fg_state.variable_groups}
  for name, data in updates.items():
    if id(name) in groups:
      flat = name.flatten(data)
      if flat.shape[-1] == 0:
        continue
      out = _set_slice(out, fg_state.vars_to_starts[name.variables[0]], flat)
    elif isinstance(name, tuple) and name in fg_state.vars_to_starts:
      data = _as_host(data)
      if data.shape[-1:] != (name[1],) or data.ndim > 2:
        raise ValueError(
            f"Expected evidence shape {(name[1],)} for variable {name}."
            f" Got incompatible shape {data.shape}."
        )
      out = _set_slice(out, fg_state.vars_to_starts[name], data)
    else:
      raise ValueError(
          "Got evidence for a variable or a VarGroup not in the FactorGraph!"
      )
  return out


class LogPotentials:
  """Log potentials of a factor graph, addressable by FactorGroup."""

  def __init__(self, fg_state: fgraph.FactorGraphState, value: Optional[np.ndarray] = None):
    self.fg_state = fg_state
    if value is None:
      value = fg_state.log_potentials
    elif value.shape != fg_state.log_potentials.shape:
      raise ValueError(
          "Expected log potentials shape"
          f" {fg_state.log_potentials.shape}. Got {value.shape}."
      )
    self.value = value

  def __getitem__(self, factor_group: fgroup.FactorGroup) -> np.ndarray:
    starts = self.fg_state.factor_group_to_potentials_starts
    if factor_group not in starts:
      raise ValueError("Invalid FactorGroup queried to access log potentials.")
    start = starts[factor_group]
    return self.value[start : start + factor_group.factor_group_log_potentials.shape[0]]

  def __setitem__(self, factor_group: fgroup.FactorGroup, data) -> None:
    self.value = update_log_potentials(self.value, {factor_group: data}, self.fg_state)


class FToVMessages:
  """Factor->variable messages of a factor graph."""

  def __init__(self, fg_state: fgraph.FactorGraphState, value: Optional[np.ndarray] = None):
    self.fg_state = fg_state
    expected = (fg_state.total_factor_num_states,)
    if value is None:
      value = np.zeros(expected)
    elif value.shape != expected:
      raise ValueError(f"Expected messages shape {expected}. Got {value.shape}.")
    self.value = value

  def __setitem__(self, variable: Tuple[int, int], data) -> None:
    """Spreads ``data`` uniformly over all messages into ``variable``."""
    self.value = update_ftov_msgs(self.value, {variable: data}, self.fg_state)


class Evidence:
  """Evidence (unary log potentials) of a factor graph."""

  def __init__(self, fg_state: fgraph.FactorGraphState, value: Optional[np.ndarray] = None):
    self.fg_state = fg_state
    expected = (fg_state.num_var_states,)
    if value is None:
      value = np.zeros(expected)
    elif value.shape != expected:
      raise ValueError(f"Expected evidence shape {expected}. Got {value.shape}.")
    self.value = value

  def __getitem__(self, variable: Tuple[int, int]) -> np.ndarray:
    start = self.fg_state.vars_to_starts[variable]
    return self.value[start : start + variable[1]]

  def __setitem__(self, name: Any, data) -> None:
    self.value = update_evidence(self.value, {name: data}, self.fg_state)


@dataclasses.dataclass(frozen=True, eq=False)
class BPState:
  """Log potentials + messages + evidence derived from one FactorGraphState."""

  log_potentials: LogPotentials
  ftov_msgs: FToVMessages
  evidence: Evidence

  def __post_init__(self):
    if not (
        self.log_potentials.fg_state
        is self.ftov_msgs.fg_state
        is self.evidence.fg_state
    ):
      raise ValueError(
          "log_potentials, ftov_msgs and evidence should be derived from the"
          " same fg_state."
      )

  @property
  def fg_state(self) -> fgraph.FactorGraphState:
    return self.log_potentials.fg_state
