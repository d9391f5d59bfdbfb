"""CPU oracle: a NumPy fp32 restatement of the reference's belief propagation path.

TEST INFRASTRUCTURE ONLY.  Nothing under pgmax_b200/ imports this module; it is
used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
--impl reference legs, always as the checker or the reported CPU baseline and
never as the product path.

The reference (pgmax 0.6.1) is pure Python on JAX; JAX is not installable in
this image, so the oracle restates, operation by operation and in fp32 (JAX's
default, x64 disabled), the jnp code of:

  run_bp / update             pgmax/infer/bp.py:85-155
  pass_var_to_fac_messages    pgmax/infer/bp.py:217-218
  normalize_and_clip_msgs     pgmax/infer/bp.py:249-259
  pass_enum_fac_to_var_...    pgmax/factor/enum.py:451-475
  pass_logical_fac_to_var_... pgmax/factor/logical.py:561-779
  pass_pool_fac_to_var_...    pgmax/factor/pool.py:328-474
  update_utils                pgmax/factor/update_utils.py:26-190
  get_beliefs / decode        pgmax/infer/inferer.py:218-222, 259-264
  get_marginals               pgmax/infer/bp.py:283-288

``x.at[idx].add(y)`` is restated as ``np.add.at`` (serial, in index order — the
order XLA:CPU's scatter uses), ``.at[idx].max`` as ``np.maximum.at``, ``.at[idx].set``
as fancy assignment (last write wins, as XLA:CPU).

Pinned against the reference's own known answers (tests/test_oracle_golden.py):
the 84 golden messages + MAP states of tests/test_pgmax.py:63-152,252-265, the
decoded states stored in benchmark/precomputed_results/ for RBMs of 24, 40, 100 and 200
units (fixtures under tests/golden/), and the reference's OR/AND/Pool-vs-Enum equivalence tests.
"""

import contextlib
import dataclasses
from typing import Any, Dict, Optional, Tuple

import numpy as np

F32 = np.float32
NEG_INF = F32(-np.inf)            # pgmax/utils/__init__.py:37
MSG_NEG_INF = F32(-1e32)          # pgmax/utils/__init__.py:26
LOG_POTENTIAL_MAX_ABS = F32(1e6)  # pgmax/utils/__init__.py:32
TEMPERATURE_STABILITY_THRE = 0.5  # pgmax/factor/logical.py:33



@contextlib.contextmanager
def precision(dtype):
  """fp64 ARBITER (not a restatement of anything the reference runs): inside
  ``with precision(np.float64)`` every array this module creates and every operation it
  performs is float64, from the same fp32 inputs.  Tests use it to tell summation-order
  noise from real differences: |fp32 path - fp64| of the device against |fp32 oracle - fp64|
  of the serial restatement itself."""
  global F32
  saved, F32 = F32, dtype
  try:
    yield
  finally:
    F32 = saved


ENUM, OR, AND, POOL = "enum", "or", "and", "pool"
FACTOR_TYPE_ORDER = (ENUM, OR, AND, POOL)  # pgmax/factor/__init__.py:37-43


@dataclasses.dataclass
class OracleGraph:
  """The static index arrays InfererContext holds (pgmax/infer/inferer.py:65-118)."""

  var_states_for_edge_states: np.ndarray
  edge_indices_for_edge_states: np.ndarray
  num_edges: int
  num_var_states: int
  msgs_range: Dict[str, Tuple[int, int]]
  potentials_range: Dict[str, Tuple[int, int]]
  inference_arguments: Dict[str, Dict[str, Any]]
  var_num_states: np.ndarray  # per variable, for decode / marginals
  # smooth dual LP-MAP only (oracle/sdlp_oracle.py): factor of every edge-state, number of factors
  factor_indices_for_edge_states: Optional[np.ndarray] = None
  num_factors: int = 0


def graph_from_context(context) -> OracleGraph:
  """OracleGraph from a pgmax_b200.infer.InfererContext (reference-format views)."""
  from pgmax_b200 import factor  # pylint: disable=g-import-not-at-top

  names = dict(zip(factor.FACTOR_TYPES, FACTOR_TYPE_ORDER))
  fg_state = context.bp_state.fg_state
  return OracleGraph(
      var_states_for_edge_states=np.asarray(context.var_states_for_edge_states),
      edge_indices_for_edge_states=np.asarray(context.edge_indices_for_edge_states),
      num_edges=context.num_edges,
      num_var_states=fg_state.num_var_states,
      msgs_range={names[ft]: r for ft, r in context.factor_type_to_msgs_range.items()},
      potentials_range={
          names[ft]: r for ft, r in context.factor_type_to_potentials_range.items()
      },
      inference_arguments={
          names[ft]: args for ft, args in context.inference_arguments.items()
      },
      var_num_states=np.concatenate(
          [vg.num_states.reshape(-1) for vg in fg_state.variable_groups]
          + [np.empty((0,), dtype=np.int64)]
      ),
      factor_indices_for_edge_states=np.asarray(context.factor_indices_for_edge_states),
      num_factors=context.num_factors,
  )


def graph_from_flat(flat) -> OracleGraph:
  """OracleGraph from a pgmax_b200._native.FlatGraph (EnumFactor blocks only): expands the
  block descriptors into the reference's per-row arrays (pgmax/factor/enum.py:364-394)."""
  edge_ns = np.asarray(flat.edge_num_states, dtype=np.int64)
  edge_vs = np.asarray(flat.edge_var_start, dtype=np.int64)
  edge_msg_start = np.concatenate([np.cumsum(edge_ns) - edge_ns, [int(edge_ns.sum())]])
  num_es = int(edge_ns.sum())
  edge_of_es = np.repeat(np.arange(edge_ns.shape[0]), edge_ns)
  vs_of_es = edge_vs[edge_of_es] + (np.arange(num_es) - edge_msg_start[edge_of_es])
  cfg_idx, cfg_es = [], []
  factor_of_edge = np.zeros((edge_ns.shape[0],), dtype=np.int64)
  factor_shift = 0
  for blk in flat.enum_blocks:
    arity = np.asarray(blk.factor_configs).shape[1]
    factor_of_edge[blk.first_edge : blk.first_edge + blk.num_factors * arity] = factor_shift + np.repeat(
        np.arange(blk.num_factors), arity)
    factor_shift += blk.num_factors
  for blk in flat.enum_blocks:
    cfg = np.asarray(blk.factor_configs, dtype=np.int64)
    K, A = cfg.shape
    off = np.concatenate([[0], np.cumsum(edge_ns[blk.first_edge : blk.first_edge + A])])
    ns, first_msg = int(off[-1]), int(edge_msg_start[blk.first_edge])
    f = np.arange(blk.num_factors)[:, None, None]
    k = np.arange(K)[None, :, None]
    cfg_idx.append(np.broadcast_to(blk.first_potential + f * K + k, (blk.num_factors, K, A)).reshape(-1))
    cfg_es.append((first_msg + f * ns + (off[:-1][None, None, :] + cfg[None])).reshape(-1))
  empty = np.zeros((0,), dtype=np.int64)
  enum_args = dict(
      factor_configs_indices=np.concatenate(cfg_idx) if cfg_idx else empty,
      factor_configs_edge_states=np.concatenate(cfg_es) if cfg_es else empty,
      num_val_configs=int(flat.num_potentials),
      num_factors=int(sum(b.num_factors for b in flat.enum_blocks)),
  )
  # OR / AND / Pool factors (FlatLogical, global message indices): their message slices follow
  # the Enum slice in type order; a factor's edges are contiguous (parents, then the child)
  enum_edges = int(sum(b.num_factors * np.asarray(b.factor_configs).shape[1] for b in flat.enum_blocks))
  msgs_range = {ENUM: (0, int(edge_msg_start[enum_edges]) if enum_edges < edge_ns.shape[0] else num_es)}
  args = {ENUM: enum_args}
  first_edge, start = enum_edges, msgs_range[ENUM][1]
  for name, ft in (("or_factors", OR), ("and_factors", AND), ("pool_factors", POOL)):
    lg = getattr(flat, name, None)
    if lg is None or lg.num_factors == 0:
      msgs_range[ft], args[ft] = (start, start), {}
      continue
    pf = np.asarray(lg.parents_factor, dtype=np.int64)
    n_edges = int(pf.shape[0]) + lg.num_factors
    end = int(edge_msg_start[first_edge + n_edges]) if first_edge + n_edges < edge_ns.shape[0] else num_es
    msgs_range[ft] = (start, end)
    counts = np.bincount(pf, minlength=lg.num_factors) + 1
    factor_of_edge[first_edge : first_edge + n_edges] = factor_shift + np.repeat(np.arange(lg.num_factors), counts)
    factor_shift += lg.num_factors
    parents = np.stack([pf, np.asarray(lg.parents_msg, dtype=np.int64) - start], axis=1)
    children = np.asarray(lg.children_msg, dtype=np.int64) - start
    if ft == POOL:
      args[ft] = dict(pool_choices_factor_indices=parents[:, 0], pool_choices_msg_indices=parents[:, 1],
                      pool_indicators_edge_states=children)
    else:
      args[ft] = dict(parents_factor_indices=parents[:, 0], parents_msg_indices=parents[:, 1],
                      children_edge_states=children, edge_states_offset=1 if ft == OR else -1)
    first_edge, start = first_edge + n_edges, end
  return OracleGraph(
      var_states_for_edge_states=vs_of_es,
      edge_indices_for_edge_states=edge_of_es,
      num_edges=int(edge_ns.shape[0]),
      num_var_states=int(np.asarray(flat.var_num_states, dtype=np.int64).sum()),
      msgs_range=msgs_range,
      potentials_range={ENUM: (0, int(flat.num_potentials)), OR: (0, 0), AND: (0, 0), POOL: (0, 0)},
      inference_arguments=args,
      var_num_states=np.asarray(flat.var_num_states, dtype=np.int64),
      factor_indices_for_edge_states=factor_of_edge[edge_of_es],
      num_factors=factor_shift,
  )


# ----------------------------------------------------------------------------
# update_utils.py
# ----------------------------------------------------------------------------
def _scatter_add(num, labels, data, fill=0.0):
  out = np.full((num,), fill, dtype=F32)
  np.add.at(out, labels, data.astype(F32, copy=False))
  return out


def _scatter_max(num, labels, data, fill=NEG_INF):
  out = np.full((num,), fill, dtype=data.dtype if data.dtype != np.float64 else F32)
  np.maximum.at(out, labels, data)
  return out


def get_maxes_and_argmaxes(data, labels, num_labels):
  """Per-label max and arg-max; ties resolve to the LARGEST index
  (pgmax/factor/update_utils.py:26-64)."""
  num_obs = data.shape[0]
  maxes = _scatter_max(num_labels, labels, data)
  only_maxes_pos = np.arange(num_obs, dtype=np.int64) - num_obs * (
      data != maxes[labels]
  ).astype(np.int64)
  argmaxes = np.full((num_labels,), np.iinfo(np.int32).min, dtype=np.int64)
  np.maximum.at(argmaxes, labels, only_maxes_pos)
  return maxes, argmaxes


def logsumexps_with_temp(data, labels, num_labels, temperature, maxes=None):
  """pgmax/factor/update_utils.py:68-98."""
  if maxes is None:
    maxes = _scatter_max(num_labels, labels, data)
  with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
    exps = np.exp((data - maxes[labels]) / temperature).astype(F32)
    return temperature * np.log(_scatter_add(num_labels, labels, exps)) + maxes


def log1mexp(x):
  """log(1 - exp(-x)), x >= 0 (pgmax/factor/update_utils.py:135-148)."""
  with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
    return np.where(
        x <= F32(np.log(2)), np.log(-np.expm1(-x)), np.log1p(-np.exp(-x))
    ).astype(F32)


def logaddexp_with_temp(data1, data2, temperature):
  """pgmax/factor/update_utils.py:151-167."""
  maxes = np.maximum(data1, data2)
  mins = np.minimum(data1, data2)
  with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
    return (temperature * np.log1p(np.exp((mins - maxes) / temperature)) + maxes).astype(F32)


def logminusexp_with_temp(data1, data2, temperature, eps=1e-30):
  """pgmax/factor/update_utils.py:170-190."""
  with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
    return np.where(
        data1 >= data2 + F32(eps),
        temperature * log1mexp((data1 - data2) / temperature) + data1,
        NEG_INF,
    ).astype(F32)


# ----------------------------------------------------------------------------
# bp.py helpers
# ----------------------------------------------------------------------------
def pass_var_to_fac_messages(ftov_msgs, evidence, var_states_for_edge_states):
  """pgmax/infer/bp.py:217-218."""
  var_sums = evidence.astype(F32).copy()
  np.add.at(var_sums, var_states_for_edge_states, ftov_msgs)
  return var_sums[var_states_for_edge_states] - ftov_msgs


def normalize_and_clip_msgs(msgs, edge_indices_for_edge_states, num_edges):
  """pgmax/infer/bp.py:249-259."""
  max_by_edges = _scatter_max(num_edges, edge_indices_for_edge_states, msgs)
  with np.errstate(invalid="ignore"):
    norm = msgs - max_by_edges[edge_indices_for_edge_states]
  return np.maximum(norm, MSG_NEG_INF)  # jnp.clip(x, MSG_NEG_INF, None)


# ----------------------------------------------------------------------------
# factor -> variable updates
# ----------------------------------------------------------------------------
def pass_enum_fac_to_var_messages(
    vtof_msgs, log_potentials, temperature, factor_configs_indices,
    factor_configs_edge_states, num_val_configs, num_factors, normalize=True):
  """pgmax/factor/enum.py:451-475."""
  del num_factors, normalize
  summary = _scatter_add(
      num_val_configs, factor_configs_indices, vtof_msgs[factor_configs_edge_states]
  ) + log_potentials
  maxes = _scatter_max(
      vtof_msgs.shape[0], factor_configs_edge_states, summary[factor_configs_indices]
  )
  if temperature == 0.0:
    ftov = maxes
  else:
    ftov = logsumexps_with_temp(
        summary[factor_configs_indices], factor_configs_edge_states,
        vtof_msgs.shape[0], temperature, maxes=maxes)
  with np.errstate(invalid="ignore"):
    return (ftov - vtof_msgs).astype(F32)


def pass_logical_fac_to_var_messages(
    vtof_msgs, log_potentials, temperature, parents_factor_indices,
    parents_msg_indices, children_edge_states, edge_states_offset, normalize=True):
  """pgmax/factor/logical.py:561-779."""
  del log_potentials
  T = temperature
  pfi = parents_factor_indices
  num_factors = children_edge_states.shape[0]
  p_rel = vtof_msgs[parents_msg_indices + edge_states_offset]
  p_oth = vtof_msgs[parents_msg_indices]
  p_diffs = p_rel - p_oth
  c_rel = vtof_msgs[children_edge_states + edge_states_offset]
  c_oth = vtof_msgs[children_edge_states]

  fsum_p_oth = _scatter_add(num_factors, pfi, p_oth)
  children_msgs_other = fsum_p_oth
  first_max, first_argmax = get_maxes_and_argmaxes(p_diffs, pfi, num_factors)
  masked = p_diffs.copy()
  masked[first_argmax] = NEG_INF
  second_max = _scatter_max(num_factors, pfi, masked)

  if T == 0.0:
    maxes_by_edge = np.maximum(p_oth, p_rel)
    fsum_maxes = _scatter_add(num_factors, pfi, maxes_by_edge)
    parents_msgs_relevant = fsum_maxes[pfi] + c_rel[pfi] - maxes_by_edge
    min0_first = np.minimum(F32(0.0), first_max)
    children_msgs_relevant = fsum_maxes + min0_first
    opt1 = c_oth[pfi] + fsum_p_oth[pfi] - p_oth
    opt2 = parents_msgs_relevant + min0_first[pfi]
    opt2[first_argmax] = parents_msgs_relevant[first_argmax] + np.minimum(F32(0.0), second_max)
    parents_msgs_other = np.maximum(opt1, opt2)
  else:
    lse_by_edge = logaddexp_with_temp(p_rel, p_oth, T)
    fsum_lse = _scatter_add(num_factors, pfi, lse_by_edge)
    fsum_lse_wo_self = fsum_lse[pfi] - lse_by_edge
    fsum_oth_wo_self = fsum_p_oth[pfi] - p_oth
    parents_msgs_relevant = c_rel[pfi] + fsum_lse_wo_self
    children_msgs_relevant = logminusexp_with_temp(fsum_lse, fsum_p_oth, T, eps=1e-4)
    if T < TEMPERATURE_STABILITY_THRE:
      children_msgs_relevant = np.maximum(
          children_msgs_relevant,
          logaddexp_with_temp(fsum_p_oth + first_max, fsum_p_oth + second_max, T),
      )
    opt1 = c_oth[pfi] + fsum_oth_wo_self
    opt2 = c_rel[pfi] + fsum_lse_wo_self
    opt3 = c_rel[pfi] + fsum_oth_wo_self
    lae12 = logaddexp_with_temp(opt1, opt2, T)
    parents_msgs_other = logminusexp_with_temp(lae12, opt3, T, eps=1e-4)
    if T < TEMPERATURE_STABILITY_THRE:
      plus_max = fsum_oth_wo_self + first_max[pfi]
      plus_max[first_argmax] = fsum_oth_wo_self[first_argmax] + second_max
      lower = logaddexp_with_temp(opt1, c_rel[pfi] + plus_max, T)
      parents_msgs_other = np.maximum(parents_msgs_other, lower)

  # Factors with a single parent (logical.py:739-757).
  num_parents = np.bincount(pfi, minlength=num_factors)
  first_elements = np.concatenate([[0], np.cumsum(num_parents)])[:-1]
  parents_msgs_relevant[first_elements] = np.where(
      num_parents == 1, c_rel, parents_msgs_relevant[first_elements])
  parents_msgs_other[first_elements] = np.where(
      num_parents == 1, c_oth, parents_msgs_other[first_elements])

  ftov = np.zeros_like(vtof_msgs)
  with np.errstate(invalid="ignore"):
    if normalize:
      ftov[parents_msg_indices + edge_states_offset] = parents_msgs_relevant - parents_msgs_other
      ftov[children_edge_states + edge_states_offset] = children_msgs_relevant - children_msgs_other
    else:
      ftov[parents_msg_indices + edge_states_offset] = parents_msgs_relevant
      ftov[parents_msg_indices] = parents_msgs_other
      ftov[children_edge_states + edge_states_offset] = children_msgs_relevant
      ftov[children_edge_states] = children_msgs_other
  return ftov.astype(F32)


def pass_pool_fac_to_var_messages(
    vtof_msgs, log_potentials, temperature, pool_choices_factor_indices,
    pool_choices_msg_indices, pool_indicators_edge_states, normalize=True):
  """pgmax/factor/pool.py:328-474."""
  del log_potentials
  T = temperature
  pfi = pool_choices_factor_indices
  num_factors = pool_indicators_edge_states.shape[0]
  choice_diffs = vtof_msgs[pool_choices_msg_indices + 1] - vtof_msgs[pool_choices_msg_indices]
  choice_zeros = vtof_msgs[pool_choices_msg_indices]
  ind_diffs = vtof_msgs[pool_indicators_edge_states + 1] - vtof_msgs[pool_indicators_edge_states]
  ind_ones = vtof_msgs[pool_indicators_edge_states + 1]

  sums_zeros = _scatter_add(num_factors, pfi, choice_zeros)
  choices_msgs_ones = sums_zeros[pfi] + ind_ones[pfi] - choice_zeros
  indicators_msgs_zeros = sums_zeros
  diffs_max, diffs_argmax = get_maxes_and_argmaxes(choice_diffs, pfi, num_factors)

  if T == 0.0:
    wo_max = choice_diffs.copy()
    wo_max[diffs_argmax] = NEG_INF
    second_max = _scatter_max(num_factors, pfi, wo_max)
    choices_msgs_diffs = np.minimum(ind_diffs, -diffs_max)[pfi]
    choices_msgs_diffs[diffs_argmax] = np.minimum(ind_diffs, -second_max)
    indicators_msgs_diffs = diffs_max
  else:
    indicators_msgs_diffs = logsumexps_with_temp(choice_diffs, pfi, num_factors, T, maxes=diffs_max)
    factor_lse = logaddexp_with_temp(indicators_msgs_diffs, -ind_diffs, T)
    choices_msgs_diffs = -logminusexp_with_temp(factor_lse[pfi], choice_diffs, T)
    replaced = choice_diffs.copy()
    replaced[diffs_argmax] = -ind_diffs
    at_argmax = -logsumexps_with_temp(replaced, pfi, num_factors, T)
    choices_msgs_diffs[diffs_argmax] = at_argmax

  # Factors with a single pool choice (pool.py:430-450).
  num_choices = np.bincount(pfi, minlength=num_factors)
  first_choices = np.concatenate([[0], np.cumsum(num_choices)])[:-1]
  choices_msgs_diffs[first_choices] = np.where(
      num_choices == 1, ind_diffs, choices_msgs_diffs[first_choices])
  choices_msgs_ones[first_choices] = np.where(
      num_choices == 1, ind_ones, choices_msgs_ones[first_choices])

  ftov = np.zeros_like(vtof_msgs)
  if normalize:
    ftov[pool_choices_msg_indices + 1] = choices_msgs_diffs
    ftov[pool_indicators_edge_states + 1] = indicators_msgs_diffs
  else:
    ftov[pool_choices_msg_indices + 1] = choices_msgs_ones
    ftov[pool_choices_msg_indices] = choices_msgs_ones - choices_msgs_diffs
    ftov[pool_indicators_edge_states + 1] = indicators_msgs_zeros + indicators_msgs_diffs
    ftov[pool_indicators_edge_states] = indicators_msgs_zeros
  return ftov.astype(F32)


FAC_TO_VAR_UPDATES = {
    ENUM: pass_enum_fac_to_var_messages,
    OR: pass_logical_fac_to_var_messages,
    AND: pass_logical_fac_to_var_messages,
    POOL: pass_pool_fac_to_var_messages,
}


# ----------------------------------------------------------------------------
# run_bp
# ----------------------------------------------------------------------------
def bp_update(graph: OracleGraph, msgs, evidence, log_potentials, damping, temperature):
  """One iteration (the ``update`` closure, pgmax/infer/bp.py:98-137)."""
  vtof = pass_var_to_fac_messages(msgs, evidence, graph.var_states_for_edge_states)
  ftov = np.zeros_like(vtof)
  for ft in FACTOR_TYPE_ORDER:
    ms, me = graph.msgs_range[ft]
    ps, pe = graph.potentials_range[ft]
    if ms != me:
      ftov[ms:me] = FAC_TO_VAR_UPDATES[ft](
          vtof_msgs=vtof[ms:me], log_potentials=log_potentials[ps:pe],
          temperature=temperature, normalize=True, **graph.inference_arguments[ft])
  d = F32(damping)
  with np.errstate(invalid="ignore", over="ignore"):
    new_msgs = d * msgs + (F32(1) - d) * ftov
    new_msgs = normalize_and_clip_msgs(
        new_msgs, graph.edge_indices_for_edge_states, graph.num_edges)
    delta = np.max(np.abs(new_msgs - msgs)) if msgs.size else F32(0)
  return new_msgs.astype(F32), F32(delta)


def run_bp(graph: OracleGraph, log_potentials, ftov_msgs, evidence, num_iters,
           damping=0.5, temperature=0.0):
  """run_with_diffs for ONE sample (pgmax/infer/bp.py:63-155): returns (msgs, deltas)."""
  temperature = float(temperature)
  T = temperature if temperature == 0.0 else F32(temperature)
  lp = np.clip(np.asarray(log_potentials, dtype=F32), -LOG_POTENTIAL_MAX_ABS, LOG_POTENTIAL_MAX_ABS)
  ev = np.asarray(evidence, dtype=F32)
  msgs = normalize_and_clip_msgs(
      np.asarray(ftov_msgs, dtype=F32), graph.edge_indices_for_edge_states, graph.num_edges)
  deltas = []
  for _ in range(max(int(num_iters), 1)):
    msgs, delta = bp_update(graph, msgs, ev, lp, damping, T)
    deltas.append(delta)
  return msgs, np.asarray(deltas, dtype=F32)


def run_bp_trajectory(graph: OracleGraph, log_potentials, ftov_msgs, evidence, num_iters,
                      damping=0.5, temperature=0.0):
  """run_bp for ONE sample that also returns the messages after every iteration:
  [num_iters, E_s] (the same updates in the same order as run_bp)."""
  temperature = float(temperature)
  T = temperature if temperature == 0.0 else F32(temperature)
  lp = np.clip(np.asarray(log_potentials, dtype=F32), -LOG_POTENTIAL_MAX_ABS, LOG_POTENTIAL_MAX_ABS)
  ev = np.asarray(evidence, dtype=F32)
  msgs = normalize_and_clip_msgs(
      np.asarray(ftov_msgs, dtype=F32), graph.edge_indices_for_edge_states, graph.num_edges)
  out = []
  for _ in range(max(int(num_iters), 1)):
    msgs, _ = bp_update(graph, msgs, ev, lp, damping, T)
    out.append(msgs)
  return np.stack(out)


def run_bp_batched(graph, log_potentials, ftov_msgs, evidence, num_iters, damping=0.5,
                   temperature=0.0):
  """vmap of run_bp over a leading axis present on any of the three arrays."""
  arrs = [np.asarray(a) for a in (log_potentials, ftov_msgs, evidence)]
  sizes = {a.shape[0] for a in arrs if a.ndim == 2}
  if not sizes:
    return run_bp(graph, *arrs, num_iters, damping, temperature)
  (batch,) = sizes
  pick = lambda a, b: a[b] if a.ndim == 2 else a
  outs = [
      run_bp(graph, pick(arrs[0], b), pick(arrs[1], b), pick(arrs[2], b), num_iters,
             damping, temperature)
      for b in range(batch)
  ]
  return np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs])


def flat_beliefs(graph: OracleGraph, ftov_msgs, evidence):
  """pgmax/infer/inferer.py:218-222 (one sample or batched)."""
  ftov_msgs, evidence = np.asarray(ftov_msgs, dtype=F32), np.asarray(evidence, dtype=F32)
  if ftov_msgs.ndim == 2 or evidence.ndim == 2:
    batch = ftov_msgs.shape[0] if ftov_msgs.ndim == 2 else evidence.shape[0]
    pick = lambda a, b: a[b] if a.ndim == 2 else a
    return np.stack(
        [flat_beliefs(graph, pick(ftov_msgs, b), pick(evidence, b)) for b in range(batch)])
  out = evidence.copy()
  np.add.at(out, graph.var_states_for_edge_states, ftov_msgs)
  return out


def decode_flat(graph: OracleGraph, beliefs):
  """Per-variable first arg-max (inferer.py:259-264), softmax marginals
  (bp.py:283-288) and the number of variables whose two best beliefs tie exactly."""
  beliefs = np.asarray(beliefs, dtype=F32)
  if beliefs.ndim == 2:
    outs = [decode_flat(graph, row) for row in beliefs]
    return tuple(np.stack([o[i] for o in outs]) for i in range(3))
  bounds = np.concatenate([[0], np.cumsum(graph.var_num_states)])
  states = np.zeros((graph.var_num_states.shape[0],), dtype=np.int32)
  marg = np.zeros_like(beliefs)
  ties = 0
  for v in range(states.shape[0]):
    x = beliefs[bounds[v] : bounds[v + 1]]
    if x.size == 0:
      continue
    states[v] = int(np.argmax(x))
    if x.size >= 2:
      top = np.sort(x)[-2:]
      ties += int(top[0] == top[1])
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
      mx = np.max(x)
      lse = np.log(np.sum(np.exp(x - mx), dtype=F32)) + mx
      marg[bounds[v] : bounds[v + 1]] = np.exp(x - lse)
  return states, marg, np.int32(ties)


# ----------------------------------------------------------------------------
# compute_energy (pgmax/infer/energy.py:53-148 and the per-type compute_energy)
# ----------------------------------------------------------------------------
def enum_energy(edge_states_one_hot_decoding, log_potentials, factor_configs_indices,
                factor_configs_edge_states, num_val_configs, num_factors):
  """pgmax/factor/enum.py:276-323."""
  decoded = np.ones((num_val_configs,), dtype=bool)
  np.multiply.at(decoded, factor_configs_indices,
                 edge_states_one_hot_decoding[factor_configs_edge_states].astype(bool))
  lp = np.asarray(log_potentials, dtype=F32)
  with np.errstate(invalid="ignore"):
    lp = np.where(np.isinf(lp) & ~decoded, -NEG_INF * np.sign(lp), lp)
  if int(decoded.sum()) != num_factors:
    return F32(np.inf)  # invalid decoding
  return F32(-np.sum(lp[decoded], dtype=F32))


def logical_energy(edge_states_one_hot_decoding, parents_factor_indices, parents_msg_indices,
                   children_edge_states, edge_states_offset, log_potentials=None):
  """pgmax/factor/logical.py:295-358."""
  del edge_states_offset, log_potentials
  num_factors = children_edge_states.shape[0]
  parents = np.ones((num_factors,), dtype=F32)
  np.multiply.at(parents, parents_factor_indices, edge_states_one_hot_decoding[parents_msg_indices])
  children = edge_states_one_hot_decoding[children_edge_states]
  return F32(np.inf) if np.any(parents != children) else F32(0.0)


def pool_energy(edge_states_one_hot_decoding, pool_choices_factor_indices, pool_choices_msg_indices,
                pool_indicators_edge_states, log_potentials=None):
  """pgmax/factor/pool.py:184-239."""
  del log_potentials
  num_factors = pool_indicators_edge_states.shape[0]
  choices = np.zeros((num_factors,), dtype=F32)
  np.add.at(choices, pool_choices_factor_indices, edge_states_one_hot_decoding[pool_choices_msg_indices + 1])
  indicators = edge_states_one_hot_decoding[pool_indicators_edge_states + 1]
  return F32(np.inf) if np.any(choices != indicators) else F32(0.0)


def compute_energy(graph: OracleGraph, log_potentials, evidence, flat_states):
  """Energy of the decoding `flat_states` ([num_vars], flat variable order) for ONE sample:
  one-hot the decoding over the var-states, -sum(one_hot * evidence), plus every factor
  type's contribution (pgmax/infer/energy.py:103-147)."""
  log_potentials = np.asarray(log_potentials, dtype=F32)
  evidence = np.asarray(evidence, dtype=F32)
  flat_states = np.asarray(flat_states, dtype=np.int64)
  bounds = np.concatenate([[0], np.cumsum(graph.var_num_states)])
  one_hot = np.zeros((graph.num_var_states,), dtype=F32)
  has = graph.var_num_states > 0
  one_hot[bounds[:-1][has] + flat_states[has]] = 1.0
  energy = F32(-np.sum(one_hot * evidence, dtype=F32))
  es_one_hot = one_hot[graph.var_states_for_edge_states]
  fns = {ENUM: enum_energy, OR: logical_energy, AND: logical_energy, POOL: pool_energy}
  with np.errstate(invalid="ignore"):
    for ft in FACTOR_TYPE_ORDER:
      ms, me = graph.msgs_range[ft]
      ps, pe = graph.potentials_range[ft]
      if ms != me:
        args = dict(graph.inference_arguments[ft])
        energy = F32(energy + fns[ft](es_one_hot[ms:me], log_potentials=log_potentials[ps:pe], **args))
  return energy
