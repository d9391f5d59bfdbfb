"""CPU oracle: the primal of the LP-MAP relaxation, solved with SciPy's HiGHS.

TEST INFRASTRUCTURE ONLY (never imported by pgmax_b200/).  Restates the LINEAR PROGRAM the
reference builds in pgmax/utils/primal_lp.py:31-178 - same variables, same constraints, same
objective - and hands it to scipy.optimize.linprog instead of cvxpy / ECOS (neither is in the
image).  The reference's own tests use exactly this program as the known answer for the smooth
dual LP-MAP solver: tests/lp/test_dual_lp.py:123,221,304,395 assert that the dual's primal upper
bound equals the LP optimum (rtol 5e-3) on tight relaxations.  An independent solver on the same
program pins oracle/sdlp_oracle.py and the device solver to a number that does not come from our
own restatement of dual_lp.py.

  variables   mu_v(s) >= 0 per variable state, sum_s mu_v(s) = 1       (primal_lp.py:63-70)
              mu_f(k) >= 0 per valid configuration of every EnumFactor (:88-96)
  objective   max  sum_v <mu_v, evidence_v> + sum_f <mu_f, log_potentials_f>   (:70,:98-100)
  EnumFactor  sum_{k: config k assigns s to v} mu_f(k) = mu_v(s)        (:103-115)
  ORFactor    child(1) <= sum parents(1);  parent(1) <= child(1)        (:118-138)
  ANDFactor   sum parents(1) <= child(1) + n - 1;  child(1) <= parent(1) (:141-162)
  PoolFactor  sum choices(1) = indicator(1)                              (:165-176)
"""

from typing import Any, Dict, Optional, Tuple

import numpy as np
import scipy.optimize
import scipy.sparse


def primal_lp_solver(fg, evidence_updates: Optional[Dict[Any, Any]] = None) -> Tuple[np.ndarray, float]:
  """(flat LP solution over the var-states in the graph's flat order, optimal objective value)."""
  from pgmax_b200 import factor  # pylint: disable=g-import-not-at-top
  from pgmax_b200.infer import bp_state as bpstate  # pylint: disable=g-import-not-at-top

  fg_state = fg.fg_state
  evidence = np.asarray(fg.bp_state.evidence.value, dtype=np.float64)
  if evidence_updates is not None:
    evidence = np.asarray(bpstate.update_evidence(evidence.astype(np.float32), evidence_updates, fg_state),
                          dtype=np.float64)
  # var-state columns: the graph's flat evidence order
  var_start = {}
  col = 0
  for vg in fg.variable_groups:
    for var in vg.variables:
      var_start[var] = col
      col += int(var[1])
  num_vs = col
  cost = [-evidence]  # linprog minimises
  eq_rows, eq_cols, eq_vals, eq_rhs = [], [], [], []
  ub_rows, ub_cols, ub_vals, ub_rhs = [], [], [], []

  def add_eq(cols, vals, rhs):
    r = len(eq_rhs)
    eq_rows.extend([r] * len(cols)); eq_cols.extend(cols); eq_vals.extend(vals); eq_rhs.append(rhs)

  def add_ub(cols, vals, rhs):
    r = len(ub_rhs)
    ub_rows.extend([r] * len(cols)); ub_cols.extend(cols); ub_vals.extend(vals); ub_rhs.append(rhs)

  for var, start in var_start.items():
    add_eq(list(range(start, start + int(var[1]))), [1.0] * int(var[1]), 1.0)
  for factor_type, groups in fg.factor_groups.items():
    for group in groups:
      if factor_type is factor.EnumFactor:
        configs = np.asarray(group.factor_configs)
        lps = np.asarray(group.log_potentials, dtype=np.float64)
        lps = np.broadcast_to(lps, (len(group.variables_for_factors), configs.shape[0]))
        for variables, lp in zip(group.variables_for_factors, lps):
          first = col
          col += configs.shape[0]
          cost.append(-lp)
          for idx, var in enumerate(variables):
            for state in range(int(var[1])):
              ks = np.flatnonzero(configs[:, idx] == state)
              add_eq([first + int(k) for k in ks] + [var_start[var] + state], [1.0] * len(ks) + [-1.0], 0.0)
      elif factor_type is factor.ORFactor:
        for variables in group.variables_for_factors:
          parents, child = variables[:-1], variables[-1]
          add_ub([var_start[child] + 1] + [var_start[p] + 1 for p in parents], [1.0] + [-1.0] * len(parents), 0.0)
          for p in parents:
            add_ub([var_start[p] + 1, var_start[child] + 1], [1.0, -1.0], 0.0)
      elif factor_type is factor.ANDFactor:
        for variables in group.variables_for_factors:
          parents, child = variables[:-1], variables[-1]
          add_ub([var_start[p] + 1 for p in parents] + [var_start[child] + 1], [1.0] * len(parents) + [-1.0],
                 float(len(parents) - 1))
          for p in parents:
            add_ub([var_start[child] + 1, var_start[p] + 1], [1.0, -1.0], 0.0)
      elif factor_type is factor.PoolFactor:
        for variables in group.variables_for_factors:
          choices, indicator = variables[:-1], variables[-1]
          add_eq([var_start[c] + 1 for c in choices] + [var_start[indicator] + 1], [1.0] * len(choices) + [-1.0], 0.0)
      else:
        raise ValueError(f"unknown factor type {factor_type}")
  c = np.concatenate(cost)
  a_eq = scipy.sparse.csr_matrix((eq_vals, (eq_rows, eq_cols)), shape=(len(eq_rhs), col))
  a_ub = scipy.sparse.csr_matrix((ub_vals, (ub_rows, ub_cols)), shape=(len(ub_rhs), col)) if ub_rhs else None
  res = scipy.optimize.linprog(c, A_ub=a_ub, b_ub=np.asarray(ub_rhs) if ub_rhs else None, A_eq=a_eq,
                               b_eq=np.asarray(eq_rhs), bounds=(0, None), method="highs")
  if res.status != 0:
    raise RuntimeError(f"linprog failed: {res.message}")
  return res.x[:num_vs], float(-res.fun)
