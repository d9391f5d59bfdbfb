"""infer.compute_energy on the device (pgx_energy) against the oracle's restatement of
pgmax/infer/energy.py and against the reference's own test cases (tests/test_energy.py)."""

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup
from pgmax_b200.infer.bp_state import BPArrays

pytestmark = pytest.mark.gpu


def _both(fg, bp, arrays, map_states):
  energy = infer.compute_energy(fg.bp_state, arrays, map_states)[0]
  energy_debug, var_e, fac_e = infer.compute_energy(fg.bp_state, arrays, map_states, debug_mode=True)
  return energy, energy_debug, var_e, fac_e


def test_energy_empty_and_single_state_vgroups():
  """tests/test_energy.py:24-72: an empty VarGroup and a single-state VarGroup; energy 0."""
  for extra in (vgroup.NDVarArray(num_states=2, shape=(0, 2)), vgroup.NDVarArray(num_states=1, shape=(1, 2))):
    variables = vgroup.NDVarArray(num_states=2, shape=(2, 2))
    fg = fgraph.FactorGraph(variable_groups=[variables, extra])
    fg.add_factors(fgroup.EnumFactorGroup(
        variables_for_factors=[[variables[0, 0], variables[0, 1]]], factor_configs=np.zeros((1, 2), int)))
    bp = infer.build_inferer(fg.bp_state, backend="bp")
    arrays = bp.init()
    map_states = infer.decode_map_states(bp.get_beliefs(arrays))
    energy, energy_debug, _, _ = _both(fg, bp, arrays, map_states)
    assert energy == 0
    assert energy == energy_debug


@pytest.mark.parametrize("all_infinite", [False, True])
def test_energy_infinite_log_potentials(all_infinite):
  """tests/test_energy.py:75-128: all but one (all) potentials -inf -> MAP [1, 1] and energy 0
  (energy +inf)."""
  variables = vgroup.NDVarArray(num_states=2, shape=(2,))
  fg = fgraph.FactorGraph(variable_groups=[variables])
  last = -np.inf if all_infinite else 0.0
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[variables[0], variables[1]]],
      log_potential_matrix=np.array([[-np.inf, -np.inf], [-np.inf, last]])))
  bp = infer.build_inferer(fg.bp_state, backend="bp")
  arrays = bp.run(bp.init(), num_iters=1, temperature=0)
  map_states = infer.decode_map_states(bp.get_beliefs(arrays))
  if not all_infinite:
    assert np.all(map_states[variables] == np.array([1, 1]))
  energy, energy_debug, _, _ = _both(fg, bp, arrays, map_states)
  assert energy == (np.inf if all_infinite else 0)
  assert energy == energy_debug


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_energy_ising_and_rbm_vs_oracle(temperature):
  """Batched decodings of an Ising grid and of an RBM (pairwise + unary EnumFactors): the
  device energy of the decoded MAP states equals the oracle's one-hot restatement; a random
  decoding too."""
  for fg, evidence in (_ising(), _rbm()):
    bp = infer.BP(fg.bp_state, temperature=temperature)
    arrays = bp.run(bp.init(evidence_updates=evidence), num_iters=8, damping=0.5)
    graph = bp_oracle.graph_from_context(bp.context)
    states, _, _ = bp.context.decode(arrays)
    rng = np.random.default_rng(0)
    for flat in (states, rng.integers(0, 2, size=states.shape).astype(np.int32)):
      got = infer.compute_energy(fg.bp_state, arrays, bp.context.unflatten_states(flat))[0]
      want = np.array([bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence[b], flat[b])
                       for b in range(flat.shape[0])])
      np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-4)


def _ising():
  fg, variables, ev = models.ising_model(n=12, batch=5)
  return fg, {variables: ev}


def _rbm():
  rs = np.random.RandomState(3)
  fg, hidden, visible = models.rbm_model(rs.normal(size=(7, 11)), rs.logistic(size=7), rs.logistic(size=11))
  rng = np.random.default_rng(1)
  return fg, {hidden: rng.gumbel(size=(4, 7, 2)).astype(np.float32),
              visible: rng.gumbel(size=(4, 11, 2)).astype(np.float32)}


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
def test_energy_logical_valid_and_violated(kind):
  """OR / AND / Pool constraints: the MAP decoding of the closed-form factors is a valid
  configuration (finite energy, equal to the oracle and to the debug-mode host loop); flipping
  one child / indicator violates its factor: +inf everywhere."""
  data = models.logical_pair(kind, 2)
  entry = data["graphs"][0]
  fg = entry[0]
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = models.init_logical(bp, entry, data)
  arrays = bp.run(arrays, num_iters=20, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  states, _, _ = bp.context.decode(arrays)
  flat = np.asarray(states)
  want = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, flat)
  got, got_debug, _, _ = _both(fg, bp, arrays, bp.context.unflatten_states(flat))
  if np.isfinite(want):
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(got_debug, want, rtol=1e-5, atol=1e-3)
  else:
    assert got == want and got_debug == want
  # every single-variable flip: device == oracle (finite or +inf alike)
  for v in range(flat.shape[0]):
    flipped = flat.copy()
    flipped[v] = 1 - flipped[v]
    want = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, flipped)
    got = infer.compute_energy(fg.bp_state, arrays, bp.context.unflatten_states(flipped))[0]
    if np.isfinite(want):
      np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-4)
    else:
      assert got == want


def test_energy_rcn_shaped_large_tables():
  """625-state variables with box configuration tables: the decoded configuration is found
  through the transposed lists; an invalid pair (outside the box) gives +inf."""
  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=5, radii=(2, 4), extra_edges=1, seed=2)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  rng = np.random.default_rng(5)
  lp = rng.normal(size=arrays.log_potentials.shape).astype(np.float32)
  arrays = BPArrays(log_potentials=lp, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  arrays = bp.run(arrays, num_iters=10, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  states, _, _ = bp.context.decode(arrays)
  flat = np.asarray(states)
  want = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, flat)
  got = infer.compute_energy(fg.bp_state, arrays, bp.context.unflatten_states(flat))[0]
  assert np.isfinite(want)
  np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-4)
  far = flat.copy()
  far[0] = (far[0] + 312) % 625  # 12 rows away on the 25 x 25 state grid: outside every box
  want = bp_oracle.compute_energy(graph, arrays.log_potentials, arrays.evidence, far)
  got = infer.compute_energy(fg.bp_state, arrays, bp.context.unflatten_states(far))[0]
  assert want == np.inf and got == np.inf
