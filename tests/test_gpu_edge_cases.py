"""Edge cases of the inference path the reference's tests touch (tests/fgraph/test_fgraph.py,
tests/test_pgmax.py, tests/test_energy.py): graphs without factors, empty and single-state
variable groups beside connected ones, unconnected variables, every small batch size."""

import warnings

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup

pytestmark = pytest.mark.gpu


def test_graph_without_factors():
  """No factor at all: run is a no-op on an empty message vector, beliefs are the evidence."""
  variables = vgroup.NDVarArray(num_states=3, shape=(4,))
  fg = fgraph.FactorGraph(variable_groups=[variables])
  bp = infer.BP(fg.bp_state, temperature=0.0)
  ev = np.random.default_rng(0).normal(size=(4, 3)).astype(np.float32)
  arrays = bp.init(evidence_updates={variables: ev})
  out, deltas = bp.run_with_diffs(arrays, num_iters=3, damping=0.5)
  assert np.asarray(out.ftov_msgs).shape == (0,)
  assert np.asarray(deltas).shape == (3,)
  beliefs = bp.get_beliefs(out)
  np.testing.assert_array_equal(beliefs[variables], ev)
  np.testing.assert_array_equal(infer.decode_map_states(beliefs)[variables], ev.argmax(-1))
  np.testing.assert_array_equal(bp.get_map_states(out)[variables], ev.argmax(-1))


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_empty_single_state_and_unconnected_variables(temperature):
  """An empty group, a single-state group and variables no factor touches share a graph with a
  connected chain: messages, beliefs, MAP and marginals equal the oracle's."""
  chain = vgroup.NDVarArray(num_states=3, shape=(5,))
  empty = vgroup.NDVarArray(num_states=2, shape=(0, 4))
  single = vgroup.NDVarArray(num_states=1, shape=(2,))
  lonely = vgroup.NDVarArray(num_states=4, shape=(3,))
  fg = fgraph.FactorGraph(variable_groups=[chain, empty, single, lonely])
  rng = np.random.default_rng(1)
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[chain[i], chain[i + 1]] for i in range(4)],
      log_potential_matrix=rng.normal(size=(4, 3, 3))))
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[single[0], chain[0]], [single[1], chain[4]]],
      factor_configs=np.array([[0, 0], [0, 2]]), log_potentials=rng.normal(size=(2, 2))))
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={chain: rng.normal(size=(5, 3)), lonely: rng.normal(size=(3, 4))})
  got, got_d = bp.run_with_diffs(arrays, num_iters=7, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 7, 0.5, temperature)
  finite = np.isfinite(want) & (want > -1e31)
  np.testing.assert_allclose(np.asarray(got.ftov_msgs)[finite], want[finite], atol=1e-5)
  assert np.all(np.asarray(got.ftov_msgs)[~finite] <= -1e31)  # states in no configuration: floor
  states, marg, _ = bp.context.decode(got, marginals=True)
  w_states, w_marg, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, got.ftov_msgs, arrays.evidence))
  np.testing.assert_array_equal(states, w_states)
  np.testing.assert_allclose(marg, w_marg, atol=1e-5)
  decoded = bp.get_map_states(got)
  assert decoded[empty].shape == (0, 4) and decoded[single].shape == (2,) and np.all(decoded[single] == 0)
  np.testing.assert_array_equal(decoded[lonely], np.asarray(arrays.evidence)[-12:].reshape(3, 4).argmax(-1))
  energy = infer.compute_energy(fg.bp_state, got, decoded)[0]
  want_e = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, states)
  if np.isfinite(want_e):
    np.testing.assert_allclose(energy, want_e, rtol=2e-6, atol=1e-4)
  else:
    assert energy == want_e


@pytest.mark.parametrize("batch", [1, 2, 3, 5, 16, 17, 31, 32, 33])
def test_every_small_batch_size_mixed_graph(batch):
  """Tile widths 1, 2, 4, 8, 16, 32 and partial last tiles on a graph that mixes pairwise, OR
  and enum factors: every sample equals its own single-sample oracle run (max-product)."""
  data = models.logical_pair("or", 1)
  entry = data["graphs"][0]
  bp = infer.BP(entry[0].bp_state, temperature=0.0)
  base = models.init_logical(bp, entry, data)
  rng = np.random.default_rng(batch)
  ev = np.asarray(base.evidence)[None] + rng.normal(size=(batch, base.evidence.shape[-1])).astype(np.float32)
  from pgmax_b200.infer.bp_state import BPArrays
  arrays = BPArrays(log_potentials=base.log_potentials, ftov_msgs=base.ftov_msgs, evidence=ev.astype(np.float32))
  got, got_d = bp.run_with_diffs(arrays, num_iters=4, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 4, 0.5, 0.0)
  assert np.asarray(got.ftov_msgs).shape == want.shape == (batch, base.ftov_msgs.shape[-1])
  np.testing.assert_allclose(got.ftov_msgs, want, atol=2e-5)
  np.testing.assert_allclose(got_d, want_d, atol=2e-5)


def test_run_bp_alias_warns_and_matches_run():
  """bp.run_bp is the deprecated alias of bp.run (pgmax/infer/bp.py:167-176)."""
  fg, variables, evidence = models.ising_model(n=6)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates={variables: evidence})
  with warnings.catch_warnings(record=True) as caught:
    warnings.simplefilter("always")
    a = bp.run_bp(arrays, num_iters=5, damping=0.5)
  assert any("deprecated" in str(w.message) for w in caught)
  b = bp.run(arrays, num_iters=5, damping=0.5)
  np.testing.assert_array_equal(a.ftov_msgs, b.ftov_msgs)
  # inputs are never written and come back as the very same objects
  assert a.log_potentials is arrays.log_potentials and a.evidence is arrays.evidence
