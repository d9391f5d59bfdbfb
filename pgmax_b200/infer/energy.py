"""Energy of a decoding (host mirror of pgmax/infer/energy.py:30-204).

``compute_energy(bp_state, bp_arrays, map_states)`` evaluates the decoding on the
device (pgx_energy: one thread per variable / factor and sample, fixed-order sums);
``debug_mode=True`` is, as in the reference, a slow per-variable / per-factor host
loop that also returns the individual energies.
"""

from typing import Any, Dict, Hashable, Tuple

import numpy as np

from pgmax_b200.infer.bp_state import BPArrays
from pgmax_b200.infer.bp_state import BPState
from pgmax_b200.infer.bp_state import Evidence
from pgmax_b200.infer.bp_state import _is_torch
from pgmax_b200.infer.inferer import DeviceBuffers
from pgmax_b200.infer.inferer import InfererContext


def get_vars_to_map_states(map_states: Dict[Hashable, Any]) -> Dict[Hashable, Any]:
  """Maps each variable of a FactorGraph to its MAP state (pgmax/infer/energy.py:30-50)."""
  vars_to_map_states = {}
  for variable_group, vg_map_states in map_states.items():
    if np.prod(variable_group.shape) == 0:  # Skip empty variable groups
      continue
    flat = variable_group.flatten(vg_map_states)
    vars_to_map_states.update(zip(variable_group.variables, list(np.array(flat))))
  return vars_to_map_states


def flatten_map_states(bp_state: BPState, map_states: Dict[Hashable, Any]) -> np.ndarray:
  """dict VarGroup -> structured states  ->  flat int32 [num_vars] (or [B, num_vars]) in the
  order of the flat evidence vector."""
  parts, batch = [], None
  for vg in bp_state.fg_state.variable_groups:
    n = int(np.prod(vg.num_states.shape))
    if n == 0:
      continue
    if vg not in map_states:
      raise ValueError(f"map_states has no entry for the variable group {vg}")
    data = map_states[vg]
    data = data.cpu().numpy() if _is_torch(data) else np.asarray(data)
    if data.shape == vg.shape:
      flat = data.reshape(-1)
    elif data.ndim == len(vg.shape) + 1 and data.shape[1:] == vg.shape:
      flat = data.reshape(data.shape[0], -1)
      batch = data.shape[0] if batch is None else batch
      if data.shape[0] != batch:
        raise ValueError("Inconsistent batch sizes in map_states")
    else:
      raise ValueError(
          f"map_states of {vg} should be of shape {vg.shape} (optionally with a leading batch "
          f"axis). Got {data.shape}."
      )
    parts.append(flat)
  if not parts:
    return np.zeros((0,), dtype=np.int32)
  if batch is not None:
    parts = [p if p.ndim == 2 else np.broadcast_to(p, (batch, p.shape[0])) for p in parts]
  return np.ascontiguousarray(np.concatenate(parts, axis=-1), dtype=np.int32)


_CONTEXTS: Dict[int, InfererContext] = {}


def _context_for(bp_state: BPState) -> InfererContext:
  ctx = _CONTEXTS.get(id(bp_state))
  if ctx is None or ctx.bp_state is not bp_state:
    ctx = InfererContext(bp_state)
    if len(_CONTEXTS) > 8:
      _CONTEXTS.clear()
    _CONTEXTS[id(bp_state)] = ctx
  return ctx


def register_context(ctx: InfererContext) -> None:
  """Lets compute_energy reuse an inferer's context (and its device plan) for ctx.bp_state."""
  if len(_CONTEXTS) > 8:
    _CONTEXTS.clear()
  _CONTEXTS[id(ctx.bp_state)] = ctx


def compute_energy(
    bp_state: BPState,
    bp_arrays: BPArrays,
    map_states: Dict[Hashable, Any],
    debug_mode=False,
) -> Tuple[Any, Any, Any]:
  """Energy of a decoding, expressed by its MAP states (the lower the better).

  Returns (energy, None, None), or with ``debug_mode`` (energy, per-variable energies,
  per-factor energies) as the reference does.  With a leading batch axis on the arrays
  or on the map states the energy is an array [B].
  """
  if debug_mode:
    return _compute_energy_debug_mode(bp_state, bp_arrays, map_states)
  import torch  # pylint: disable=g-import-not-at-top

  ctx = _context_for(bp_state)
  flat_states = flatten_map_states(bp_state, map_states)
  buf = DeviceBuffers(bp_arrays, ctx._device())  # pylint: disable=protected-access
  plan = ctx.plan
  states_batch = flat_states.shape[0] if flat_states.ndim == 2 else None
  sizes = {s for s in (buf.batch, states_batch) if s is not None}
  if len(sizes) > 1:
    raise ValueError(f"Inconsistent batch sizes: {sorted(sizes)}")
  batch = sizes.pop() if sizes else 1
  states = torch.from_numpy(flat_states).to(buf.device)
  out = torch.empty((batch,), dtype=torch.float32, device=buf.device)
  stream = torch.cuda.current_stream(buf.device).cuda_stream
  plan.energy(stream, batch, buf.lp.data_ptr(), buf.lp.ndim == 2, buf.ev.data_ptr(),
              buf.ev.ndim == 2, states.data_ptr(), states.ndim == 2, out.data_ptr())
  energies = out.cpu().numpy()
  if buf.batch is None and states_batch is None:
    return float(energies[0]), None, None
  return energies, None, None


def _compute_energy_debug_mode(
    bp_state: BPState,
    bp_arrays: BPArrays,
    map_states: Dict[Hashable, Any],
) -> Tuple[float, Any, Any]:
  """Host loop over variables and factors (pgmax/infer/energy.py:151-204)."""
  print("Computing the energy of a decoding in debug mode is slow...")
  energy = 0.0
  vars_energies = {}
  factors_energies = {}
  vars_to_map_states = get_vars_to_map_states(map_states)
  ev = bp_arrays.evidence
  ev = ev.cpu().numpy() if _is_torch(ev) else np.asarray(ev)
  evidence = Evidence(bp_state.fg_state, value=np.array(ev))
  for variable_group in bp_state.fg_state.variable_groups:
    for var in variable_group.variables:
      var_decoded_state = int(vars_to_map_states[var])
      var_energy = -float(evidence[var][var_decoded_state])
      vars_energies[var] = var_energy
      energy += var_energy
  for factor_group in bp_state.fg_state.factor_group_to_potentials_starts:
    factor_configs = factor_group.factor_configs
    for this_factor in factor_group.factors:
      this_factor_variables = this_factor.variables
      factor_energy = factor_group.factor_type.compute_factor_energy(
          variables=this_factor_variables,
          vars_to_map_states=vars_to_map_states,
          factor_configs=factor_configs,
          log_potentials=this_factor.log_potentials,
      )
      energy += factor_energy
      factors_energies[frozenset(this_factor_variables)] = factor_energy
  return energy, vars_energies, factors_energies
