"""CPU oracle of the smooth dual LP-MAP solver: a NumPy fp32 restatement of
pgmax/infer/dual_lp.py:60-463 (pgmax 0.6.1).

TEST INFRASTRUCTURE ONLY (see oracle/bp_oracle.py): nothing under pgmax_b200/
imports this module.

  smooth_dual_objval_and_grad        pgmax/infer/dual_lp.py:67-237
  softmax_and_logsumexps_with_temp   pgmax/factor/update_utils.py:102-131
  run_with_objvals                   pgmax/infer/dual_lp.py:239-324
  get_primal_upper_bound             pgmax/infer/dual_lp.py:366-378

The factor -> variable updates are the ones of oracle/bp_oracle.py with
``normalize=False`` and the UNCLIPPED potentials, on ``vtof = -ftov_msgs``
(dual_lp.py:124-140).  Scalars follow JAX's weak typing: ``lr`` is a Python float
multiplied into fp32 arrays, ``it`` an int32 promoted to fp32.

Pinning: the reference stores no SDLP outputs; the restatement is pinned through the
reference's own property tests (tests/lp/test_dual_lp.py, tests/lp/test_bp_for_lp.py):
the dual upper bound meets the energy of the decoded primal when the relaxation is
tight, the objective is non-increasing for lr <= logsumexp_temp, and the closed-form
gradient equals a finite-difference gradient of the objective (tests/test_oracle_sdlp.py).
"""

import numpy as np

from oracle import bp_oracle
from oracle.bp_oracle import F32
from oracle.bp_oracle import NEG_INF


def softmax_and_logsumexps_with_temp(data, labels, num_labels, temperature):
  """pgmax/factor/update_utils.py:102-131."""
  T = F32(temperature)
  maxes = bp_oracle._scatter_max(num_labels, labels, data)
  with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
    exp_data = (T * np.exp((data - maxes[labels]) / T)).astype(F32)
    sumexp = bp_oracle._scatter_add(num_labels, labels, exp_data)
    logsumexp = (maxes + T * np.log(sumexp / T)).astype(F32)
    softmax = (exp_data / sumexp[labels]).astype(F32)
  return softmax, logsumexp


def bp_updates(graph, ftov_msgs, log_potentials, temperature):
  """The factor -> variable updates on vtof = -ftov_msgs, un-normalised (dual_lp.py:119-140)."""
  T = temperature if temperature == 0.0 else F32(temperature)
  out = np.zeros_like(ftov_msgs)
  for ft in bp_oracle.FACTOR_TYPE_ORDER:
    ms, me = graph.msgs_range[ft]
    ps, pe = graph.potentials_range[ft]
    if ms != me:
      out[ms:me] = bp_oracle.FAC_TO_VAR_UPDATES[ft](
          vtof_msgs=-ftov_msgs[ms:me], log_potentials=log_potentials[ps:pe],
          temperature=T, normalize=False, **graph.inference_arguments[ft])
  return out


def smooth_dual_objval_and_grad(graph, ftov_msgs, log_potentials, evidence, logsumexp_temp):
  """(objval, grad, bp_updates, per-edge logsumexp / max) for ONE sample (dual_lp.py:67-237)."""
  ftov_msgs = np.asarray(ftov_msgs, dtype=F32)
  log_potentials = np.asarray(log_potentials, dtype=F32)
  evidence = np.asarray(evidence, dtype=F32)
  vs = graph.var_states_for_edge_states
  edge_of_es = graph.edge_indices_for_edge_states
  num_vars = graph.var_num_states.shape[0]
  evidence_to_vars = np.repeat(np.arange(num_vars), graph.var_num_states)

  var_sums = evidence.copy()
  np.add.at(var_sums, vs, ftov_msgs)
  updates = bp_updates(graph, ftov_msgs, log_potentials, float(logsumexp_temp))
  with np.errstate(invalid="ignore"):
    outgoing = (updates - ftov_msgs).astype(F32)

  def per_factor_max(per_edge):
    out = np.full((graph.num_factors,), NEG_INF, dtype=F32)
    np.maximum.at(out, graph.factor_indices_for_edge_states, per_edge[edge_of_es])
    return out

  if logsumexp_temp == 0.0:
    maxes_vars, argmaxes_vars = bp_oracle.get_maxes_and_argmaxes(var_sums, evidence_to_vars, num_vars)
    plus = np.zeros((evidence.shape[0],), dtype=F32)
    plus[argmaxes_vars[argmaxes_vars >= 0]] = 1.0
    maxes_edges, argmaxes_edges = bp_oracle.get_maxes_and_argmaxes(outgoing, edge_of_es, graph.num_edges)
    minus = np.zeros((ftov_msgs.shape[0],), dtype=F32)
    minus[argmaxes_edges[argmaxes_edges >= 0]] = -1.0
    grad = plus[vs] + minus
    objval = F32(np.sum(maxes_vars, dtype=F32) + np.sum(per_factor_max(maxes_edges), dtype=F32))
    return objval, grad.astype(F32), updates, maxes_edges

  softmax_vars, lse_vars = softmax_and_logsumexps_with_temp(var_sums, evidence_to_vars, num_vars, logsumexp_temp)
  softmax_edges, lse_edges = softmax_and_logsumexps_with_temp(outgoing, edge_of_es, graph.num_edges, logsumexp_temp)
  grad = (softmax_vars[vs] - softmax_edges).astype(F32)
  objval = F32(np.sum(lse_vars, dtype=F32) + np.sum(per_factor_max(lse_edges), dtype=F32))
  return objval, grad, updates, lse_edges


def run_with_objvals(graph, log_potentials, ftov_msgs, evidence, logsumexp_temp, num_iters, lr=None):
  """(ftov_msgs, objvals[num_iters]) for ONE sample (dual_lp.py:239-324)."""
  if logsumexp_temp < 0.0 or logsumexp_temp > 1.0:
    raise ValueError(
        "The log sum-exp temperature of the Dual LP-MAP solver has to be between 0.0 and 1.0")
  if logsumexp_temp != 0.0 and lr is not None and lr > logsumexp_temp:
    raise ValueError(
        "For gradient descent, the learning rate must be smaller than the log sum-exp temperature.")
  if lr is None:
    lr = logsumexp_temp if logsumexp_temp != 0.0 else 0.01
  msgs = np.asarray(ftov_msgs, dtype=F32).copy()
  eta = msgs.copy()
  objvals = []
  for it in range(int(num_iters)):
    objval, grad, _, _ = smooth_dual_objval_and_grad(graph, msgs, log_potentials, evidence, logsumexp_temp)
    step, momentum = step_scalars(it, lr, logsumexp_temp)
    with np.errstate(invalid="ignore", over="ignore"):
      new_eta = (msgs - step * grad).astype(F32)
      msgs = (new_eta + momentum * (new_eta - eta)).astype(F32)
    eta = new_eta
    objvals.append(objval)
  return msgs, np.asarray(objvals, dtype=F32)


def step_scalars(it, lr, logsumexp_temp):
  """fp32 step size and Nesterov momentum of iteration ``it`` (dual_lp.py:293-309)."""
  itf = F32(it)
  step = F32(lr) if logsumexp_temp > 0 else F32(F32(lr) / np.sqrt(itf + F32(1.0), dtype=F32))
  momentum = F32((itf + F32(1.0)) / (itf + F32(4.0)))
  return step, momentum
