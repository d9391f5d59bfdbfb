"""Pins the oracle's restatement of the smooth dual LP-MAP solver (pgmax/infer/dual_lp.py)
through the reference's own property tests (no GPU, no stored JAX outputs exist):

  * tests/lp/test_dual_lp.py:30-135 - fully connected 3-state Ising model with a tight
    relaxation: the dual upper bound meets the energy of the decoded primal (rtol 5e-3),
    for accelerated gradient descent (T = 1e-3) and subgradient descent (T = 0);
  * tests/lp/test_dual_lp.py:139-235 - line sparsification with ORFactors: (L + 3) // 3
    top variables switch on, bounds meet;
  * tests/lp/test_dual_lp.py:231-306 - ANDFactors finding the all-ones rows of a matrix;
    tests/lp/test_dual_lp.py:309-402 - a hierarchy of PoolFactors (one variable per layer on);
  * tests/lp/test_bp_for_lp.py:28-391 - the BP updates at T = 1e-3 / 0.01 are within T of the
    max-product updates, and the per-edge max / logsumexp is the same at every edge of a factor;
  * the closed-form gradient (dual_lp.py:213-217) equals a central finite difference of the
    objective on graphs mixing Enum and OR / AND / Pool factors.
"""

import os
import sys

import numpy as np
import pytest

import models
from oracle import bp_oracle
from oracle import sdlp_oracle
from pgmax_b200 import infer

RTOL = 5e-3  # tests/lp/test_dual_lp.py:27
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdlp_lp.npz"))


def _check_lp(name, graph, upper, lower, states):
  """The reference's known answer (tests/lp/test_dual_lp.py:123,221,304,395: the dual's primal
  upper bound equals the LP optimum, rtol 5e-3) against the committed optimum of an independent
  LP solver (tests/golden/sdlp_lp.npz, SciPy HiGHS on the program of pgmax/utils/primal_lp.py);
  the relaxations are tight, so the decoding must also BE the LP's (integral) solution."""
  objval = float(GOLD[f"{name}_objval"])
  assert np.isclose(objval, upper, rtol=RTOL), (objval, upper)
  assert np.isclose(objval, lower, rtol=RTOL), (objval, lower)
  bounds = np.concatenate([[0], np.cumsum(graph.var_num_states)])
  lp_states = np.array([int(np.argmax(GOLD[f"{name}_solution"][bounds[v] : bounds[v + 1]]))
                        for v in range(len(bounds) - 1)])
  np.testing.assert_array_equal(states, lp_states)


def test_lp_fixture_is_what_the_solver_returns():
  """tests/golden/sdlp_lp.npz == a fresh HiGHS solve of oracle/primal_lp_oracle.py's program."""
  from oracle import primal_lp_oracle
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
  import make_sdlp_golden
  for name, fg, evidence in make_sdlp_golden.cases():
    _, objval = primal_lp_oracle.primal_lp_solver(fg, evidence)
    assert np.isclose(objval, float(GOLD[f"{name}_objval"]), rtol=1e-9, atol=1e-9), name


def _bounds(graph, arrays, msgs):
  upper, _, _, _ = sdlp_oracle.smooth_dual_objval_and_grad(
      graph, msgs, arrays.log_potentials, arrays.evidence, 0.0)
  beliefs = bp_oracle.flat_beliefs(graph, msgs, arrays.evidence)
  states, _, _ = bp_oracle.decode_flat(graph, beliefs)
  lower = -bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, states)
  return upper, lower, states


@pytest.mark.parametrize("seed,temp", [(0, 1e-3), (1, 0.0)])
def test_dual_bounds_meet_on_tight_ising(seed, temp):
  fg, variables = models.sdlp_ising_model(seed=seed)
  rng = np.random.RandomState(seed)
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates={variables: rng.gumbel(size=(4, 4, 3))})
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, objvals = sdlp_oracle.run_with_objvals(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, temp, 3000)
  upper, lower, states = _bounds(graph, arrays, msgs)
  assert np.isclose(lower, upper, rtol=RTOL)
  assert objvals[-1] <= objvals[0]
  _check_lp(f"ising_{seed}", graph, upper, lower, states)


def test_line_sparsification_with_or_factors():
  fg, top, bottom, evidence = models.sdlp_line_model(seed=0)
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates=evidence)
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, _ = sdlp_oracle.run_with_objvals(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1e-3, 5000)
  upper, lower, states = _bounds(graph, arrays, msgs)
  assert states[:20].sum() == (20 + 3) // 3
  assert np.isclose(lower, upper, rtol=RTOL)
  _check_lp("line_0", graph, upper, lower, states)


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
def test_gradient_is_the_finite_difference_of_the_objective(kind):
  data = models.logical_pair(kind, seed=3)
  entry = data["graphs"][0]  # half `kind` factors, half EnumFactors
  bp = infer.BP(entry[0].bp_state)
  arrays = models.init_logical(bp, entry, data)
  graph = bp_oracle.graph_from_context(bp.context)
  temp, h = 1.0, 2e-2
  msgs = np.asarray(arrays.ftov_msgs, dtype=np.float32)
  _, grad, _, _ = sdlp_oracle.smooth_dual_objval_and_grad(
      graph, msgs, arrays.log_potentials, arrays.evidence, temp)
  rng = np.random.RandomState(0)
  for e in rng.choice(msgs.shape[0], size=12, replace=False):
    vals = []
    for sign in (1.0, -1.0):
      moved = msgs.copy()
      moved[e] += sign * h
      vals.append(float(sdlp_oracle.smooth_dual_objval_and_grad(
          graph, moved, arrays.log_potentials, arrays.evidence, temp)[0]))
    assert abs((vals[0] - vals[1]) / (2 * h) - grad[e]) < 2e-2, (e, vals, grad[e])


def test_argument_checks():
  fg, _ = models.sdlp_ising_model(seed=0)
  bp = infer.BP(fg.bp_state)
  arrays = bp.init()
  graph = bp_oracle.graph_from_context(bp.context)
  with pytest.raises(ValueError, match="has to be between 0.0 and 1.0"):
    sdlp_oracle.run_with_objvals(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1.01, 1)
  with pytest.raises(ValueError, match="learning rate must be smaller"):
    sdlp_oracle.run_with_objvals(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 0.01, 1, lr=0.1)


@pytest.mark.parametrize("temperature", [1e-3, 0.01])
def test_low_temperature_updates_and_per_factor_consistency(temperature):
  """tests/lp/test_bp_for_lp.py:28-391 on the oracle (EnumFactors, ORFactors, PoolFactors):
  pins the normalize=False branches of the closed-form updates at T > 0 to their T = 0 limit."""
  for name, fg, evidence, ftype in models.lp_bp_cases():
    bp = infer.BP(fg.bp_state)
    rng = np.random.RandomState(11)
    arrays = bp.init(evidence_updates=evidence,
                     ftov_msgs_updates={ftype: rng.normal(size=fg.bp_state.ftov_msgs.value.shape)})
    graph = bp_oracle.graph_from_context(bp.context)

    def get_bp_updates(temp):
      _, _, updates, edge_vals = sdlp_oracle.smooth_dual_objval_and_grad(
          graph, arrays.ftov_msgs, arrays.log_potentials, arrays.evidence, temp)
      return updates, edge_vals

    models.check_lp_bp_properties(get_bp_updates, bp.context, temperature)


@pytest.mark.parametrize("seed", [0, 1])
def test_and_factors_rows_of_ones(seed):
  """tests/lp/test_dual_lp.py:231-306 on the oracle: the decoded `all_ones` variables are the rows
  of the observed matrix that are all ones, and the bounds meet."""
  fg, matrix, all_ones, evidence, truth = models.sdlp_and_model(seed=seed)
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates=evidence)
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, _ = sdlp_oracle.run_with_objvals(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1e-3, 5000)
  upper, lower, states = _bounds(graph, arrays, msgs)
  np.testing.assert_array_equal(states[50:], truth)
  assert np.isclose(lower, upper, rtol=RTOL)
  _check_lp(f"and_{seed}", graph, upper, lower, states)


@pytest.mark.parametrize("seed", [0, 1])
def test_pool_factor_hierarchy(seed):
  """tests/lp/test_dual_lp.py:309-402 on the oracle: with the root forced on, exactly one variable
  per layer switches on, and the bounds meet."""
  fg, variables = models.sdlp_pool_model()
  updates = np.random.RandomState(seed).gumbel(size=(variables.shape[0], 2))
  updates[0, 1] = 1_000
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates={variables: updates})
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, _ = sdlp_oracle.run_with_objvals(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1e-3, 5000)
  upper, lower, states = _bounds(graph, arrays, msgs)
  assert states.sum() == 4  # n_layers
  assert np.isclose(lower, upper, rtol=RTOL)
  _check_lp(f"pool_{seed}", graph, upper, lower, states)
