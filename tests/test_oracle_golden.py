"""Pins the CPU oracle to the reference's own known answers (no GPU needed).

(1) tests/test_pgmax.py:63-152,252-265 — 84 golden messages (atol 1e-6, the
    reference's own tolerance at :418) and 12 golden MAP states after 100
    max-product iterations on the 3x3 cut model.
(2) benchmark/precomputed_results n_units_24 — the reference's decoded RBM states
    after 20 and 200 max-product iterations (identical on its CPU and GPU
    back-ends), harness benchmark/rbm_lib.py:135-214.
"""

import os

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import infer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cut_model_golden_messages_and_map():
  gold = np.load(os.path.join(GOLDEN, "e2e_sanity.npz"))
  fg, bp_state, grid_vars, additional_vars = models.cut_model()
  bp = infer.BP(bp_state, temperature=0.0)
  graph = bp_oracle.graph_from_context(bp.context)
  arrays = bp.init()
  msgs, deltas = bp_oracle.run_bp(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
      num_iters=100, damping=0.5, temperature=0.0)
  assert msgs.shape == (84,) and deltas.shape == (100,)
  # The reference's criterion is jnp.allclose(..., atol=1e-06), i.e. rtol = 1e-5 (default).
  np.testing.assert_allclose(msgs, gold["true_final_msgs_output"], atol=1e-6, rtol=1e-5)
  assert np.max(np.abs(msgs - gold["true_final_msgs_output"])) < 2e-6

  beliefs = infer.inferer.unflatten_beliefs(
      bp_oracle.flat_beliefs(graph, msgs, arrays.evidence), fg.variable_groups)
  decoded = infer.decode_map_states(beliefs)
  groups = {"grid_vars": grid_vars, "additional_vars": additional_vars}
  for name, index, state in zip(gold["map_groups"], gold["map_indices"], gold["map_states"]):
    assert decoded[groups[str(name)]][tuple(index)] == state


@pytest.mark.parametrize("num_iters", [20, 200])
def test_rbm24_decoded_states_match_reference(num_iters):
  gold = np.load(os.path.join(GOLDEN, "rbm24.npz"))
  # The reference's CPU and GPU back-ends agree on every 24-unit RBM.
  np.testing.assert_array_equal(gold[f"hidden_cpu_{num_iters}"], gold[f"hidden_gpu_{num_iters}"])
  mismatches = 0
  for idx in range(0, 50, 5):
    W, bh, bv = gold["W"][idx], gold["bh"][idx], gold["bv"][idx]
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    graph = bp_oracle.graph_from_context(bp.context)
    arrays = bp.init()
    msgs, _ = bp_oracle.run_bp(
        graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
        num_iters=num_iters, damping=0.5, temperature=0.0)
    states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, msgs, arrays.evidence))
    pred_h, pred_v = states[: bh.shape[0]], states[bh.shape[0] :]
    same = np.array_equal(pred_h, gold[f"hidden_cpu_{num_iters}"][idx]) and np.array_equal(
        pred_v, gold[f"visible_cpu_{num_iters}"][idx])
    energy = models.rbm_energy(pred_h, pred_v, W, bh, bv)
    np.testing.assert_allclose(energy, gold[f"energy_cpu_{num_iters}"][idx], rtol=1e-5, atol=1e-5) if same else None
    mismatches += int(not same)
  assert mismatches == 0


def test_oracle_energy_matches_reference_rbm_energies():
  """The oracle's restatement of infer.compute_energy (pgmax/infer/energy.py:53-148) on the
  decoded states the reference stored, against the energies the reference stored beside them
  (benchmark/rbm_lib.py:35-61: -(h W v + bh h + bv v), i.e. the PGMax energy with zero evidence)."""
  gold = np.load(os.path.join(GOLDEN, "rbm24.npz"))
  for idx in range(0, 50, 7):
    W, bh, bv = gold["W"][idx], gold["bh"][idx], gold["bv"][idx]
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    graph = bp_oracle.graph_from_context(bp.context)
    arrays = bp.init()
    for num_iters in (20, 200):
      flat = np.concatenate([gold[f"hidden_cpu_{num_iters}"][idx], gold[f"visible_cpu_{num_iters}"][idx]])
      energy = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, flat)
      np.testing.assert_allclose(energy, gold[f"energy_cpu_{num_iters}"][idx], rtol=1e-5, atol=1e-4)


def test_oracle_energy_reference_test_cases():
  """tests/test_energy.py:75-128 on the oracle: all but one potential -inf -> energy 0 for the
  decoding [1, 1]; an invalid EnumFactor decoding -> +inf."""
  from pgmax_b200 import fgraph, fgroup, vgroup
  variables = vgroup.NDVarArray(num_states=2, shape=(2,))
  fg = fgraph.FactorGraph(variable_groups=[variables])
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[variables[0], variables[1]]],
      log_potential_matrix=np.array([[-np.inf, -np.inf], [-np.inf, 0.0]])))
  bp = infer.BP(fg.bp_state, temperature=0.0)
  graph = bp_oracle.graph_from_context(bp.context)
  arrays = bp.init()
  assert bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, [1, 1]) == 0
  assert bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, [0, 1]) == np.inf
  fg2 = fgraph.FactorGraph(variable_groups=[variables])
  fg2.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[variables[0], variables[1]]], factor_configs=np.zeros((1, 2), int)))
  bp2 = infer.BP(fg2.bp_state, temperature=0.0)
  graph2 = bp_oracle.graph_from_context(bp2.context)
  arrays2 = bp2.init()
  assert bp_oracle.compute_energy(graph2, arrays2.log_potentials, arrays2.evidence, [0, 0]) == 0
  assert bp_oracle.compute_energy(graph2, arrays2.log_potentials, arrays2.evidence, [1, 0]) == np.inf


def test_oracle_divergent_system_decodes_correctly():
  """tests/test_clipping.py:24-58 on the oracle: 45 equality factors over 10 binary variables,
  damping 0, 100 max-product iterations - the messages run into the -1e32 clip and the
  decoding must still be all ones (the MSG_NEG_INF / NEG_INF split, utils/__init__.py:26-37)."""
  from pgmax_b200 import fgraph, fgroup, infer, vgroup

  pg_vars = vgroup.NDVarArray(num_states=2, shape=(10,))
  pairs = [[pg_vars[i], pg_vars[j]] for i in range(10) for j in range(i + 1, 10)]
  fg = fgraph.FactorGraph(pg_vars)
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=pairs, factor_configs=np.array([[0, 0], [1, 1]])))
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates={pg_vars: np.transpose([np.zeros(10), np.ones(10)])})
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 100, 0.0, 0.0)
  assert np.isfinite(msgs).all() and msgs.min() >= -1e32
  states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, msgs, arrays.evidence))
  assert np.all(states == 1)


@pytest.mark.parametrize("n_units", [40, 100, 200])
def test_rbm_large_decoded_states_match_reference(n_units):
  """benchmark/precomputed_results/n_units_{40,100,200}: decoded states and energies after 20
  max-product iterations (the horizon on which the reference's CPU and GPU back-ends agree,
  SURVEY.md §8c) - the oracle reproduces them exactly."""
  gold = np.load(os.path.join(GOLDEN, "rbm_large.npz"))
  for idx in range(4):
    W, bh, bv = (gold[f"{k}_{n_units}_{idx}"] for k in ("W", "bh", "bv"))
    np.testing.assert_array_equal(gold[f"hidden_cpu_{n_units}_{idx}"], gold[f"hidden_gpu_{n_units}_{idx}"])
    np.testing.assert_array_equal(gold[f"visible_cpu_{n_units}_{idx}"], gold[f"visible_gpu_{n_units}_{idx}"])
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    graph = bp_oracle.graph_from_context(bp.context)
    arrays = bp.init()
    msgs, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                               num_iters=20, damping=0.5, temperature=0.0)
    states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, msgs, arrays.evidence))
    pred_h, pred_v = states[: bh.shape[0]], states[bh.shape[0] :]
    np.testing.assert_array_equal(pred_h, gold[f"hidden_cpu_{n_units}_{idx}"])
    np.testing.assert_array_equal(pred_v, gold[f"visible_cpu_{n_units}_{idx}"])
    np.testing.assert_allclose(models.rbm_energy(pred_h, pred_v, W, bh, bv), gold[f"energy_cpu_{n_units}_{idx}"],
                               rtol=1e-5, atol=1e-4)
