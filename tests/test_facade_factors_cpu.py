"""CPU tests of the host mirror's factor layer, restating the reference's
tests/factor/test_factor.py:24-222 (construction checks of Enum / Logical / Pool factors and
their wirings) and tests/fgroup/test_wiring.py:29-269 (a graph built from one factor group, from
many small groups and from individual factors compiles to the same flat arrays).

The mirror's Wiring classes take the per-edge tables (first var-state, number of states, factor id)
instead of the reference's expanded var_states_for_edges rows; the reference-format rows are
derived views (Wiring.var_states_for_edges) and are what is compared here."""

import re

import numpy as np
import pytest

from pgmax_b200 import factor, fgraph, fgroup, infer, vgroup


def test_enumeration_factor_checks():
  """tests/factor/test_factor.py:24-100."""
  variables = vgroup.NDVarArray(num_states=3, shape=(1,))
  with pytest.raises(NotImplementedError, match="Please implement compile_wiring in for your factor"):
    factor.Factor(variables=[variables[0]], log_potentials=np.array([0.0]))
  with pytest.raises(ValueError, match="Configurations should be integers. Got"):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([[1.0]]), log_potentials=np.array([0.0]))
  with pytest.raises(ValueError, match="Potential should be floats. Got"):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([[1]]), log_potentials=np.array([0]))
  with pytest.raises(ValueError, match="factor_configs should be a 2D array"):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([1]), log_potentials=np.array([0.0]))
  with pytest.raises(ValueError, match=re.escape("Number of variables 1 doesn't match given configurations (1, 2)")):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([[1, 2]]), log_potentials=np.array([0.0]))
  with pytest.raises(ValueError, match=re.escape("Expected log potentials of shape (1,)")):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([[1]]), log_potentials=np.array([0.0, 1.0]))
  with pytest.raises(ValueError, match="Invalid configurations for given variables"):
    factor.EnumFactor(variables=[variables[0]], factor_configs=np.array([[10]]), log_potentials=np.array([0.0]))
  with pytest.raises(ValueError, match="list_var_states_for_edges cannot be None"):
    factor.concatenate_var_states_for_edges(None)
  with pytest.raises(ValueError, match="var_states_for_edges cannot be None"):
    factor.concatenate_var_states_for_edges([None])


def _one_factor_edge_tables(num_parents):
  """Per-edge tables of ONE logical / pool factor over binary variables 0 .. num_parents."""
  n = num_parents + 1
  return np.arange(0, 2 * n, 2), np.full((n,), 2), np.zeros((n,), dtype=np.int64)


def test_logical_factor_checks():
  """tests/factor/test_factor.py:103-164."""
  child = vgroup.NDVarArray(num_states=2, shape=(1,))[0]
  wrong_parent = vgroup.NDVarArray(num_states=3, shape=(1,))[0]
  parent = vgroup.NDVarArray(num_states=2, shape=(1,))[0]
  with pytest.raises(ValueError, match="A LogicalFactor requires at least one parent variable and one child variable"):
    factor.logical.LogicalFactor(variables=(child,))
  with pytest.raises(ValueError, match="All the variables in a LogicalFactor should be binary"):
    factor.logical.LogicalFactor(variables=(wrong_parent, child))
  logical_factor = factor.logical.LogicalFactor(variables=(parent, child))
  num_parents = len(logical_factor.variables) - 1
  parents_edge_states = np.vstack([np.zeros(num_parents, dtype=int), np.arange(0, 2 * num_parents, 2, dtype=int)]).T
  child_edge_state = np.array([2 * num_parents], dtype=int)
  tables = _one_factor_edge_tables(num_parents)

  with pytest.raises(ValueError, match="The highest LogicalFactor index must be 0"):
    factor.logical.LogicalWiring(*tables, parents_edge_states=parents_edge_states + np.array([[1, 0]]),
                                 children_edge_states=child_edge_state,
                                 edge_states_offset=1).get_inference_arguments()
  two = np.vstack([parents_edge_states, parents_edge_states]) + np.array([[0, 0], [1, 0]])
  with pytest.raises(ValueError, match="The LogicalWiring must have 1 different LogicalFactor indices"):
    factor.logical.LogicalWiring(*tables, parents_edge_states=two, children_edge_states=child_edge_state,
                                 edge_states_offset=1).get_inference_arguments()
  with pytest.raises(ValueError, match=re.escape(
      "The LogicalWiring's edge_states_offset must be 1 (for OR) and -1 (for AND), but is 0")):
    factor.logical.LogicalWiring(*tables, parents_edge_states=parents_edge_states,
                                 children_edge_states=child_edge_state,
                                 edge_states_offset=0).get_inference_arguments()
  args = factor.logical.LogicalWiring(*tables, parents_edge_states=parents_edge_states,
                                      children_edge_states=child_edge_state,
                                      edge_states_offset=1).get_inference_arguments()
  assert set(args) >= {"parents_factor_indices", "parents_msg_indices", "children_edge_states", "edge_states_offset"}


def test_pool_factor_checks():
  """tests/factor/test_factor.py:167-222."""
  pool_choice = vgroup.NDVarArray(num_states=2, shape=(1,))[0]
  wrong_pool_indicator = vgroup.NDVarArray(num_states=3, shape=(1,))[0]
  pool_indicator = vgroup.NDVarArray(num_states=2, shape=(1,))[0]
  with pytest.raises(ValueError, match="A PoolFactor requires at least one pool choice and one pool indicator."):
    factor.pool.PoolFactor(variables=(pool_choice,))
  with pytest.raises(ValueError, match="All the variables in a PoolFactor should all be binary"):
    factor.pool.PoolFactor(variables=(wrong_pool_indicator, pool_choice))
  pool_factor = factor.pool.PoolFactor(variables=(pool_indicator, pool_choice))
  num_children = len(pool_factor.variables) - 1
  choices = np.vstack([np.zeros(num_children, dtype=int), np.arange(0, 2 * num_children, 2, dtype=int)]).T
  indicator = np.array([2 * num_children], dtype=int)
  tables = _one_factor_edge_tables(num_children)
  with pytest.raises(ValueError, match="The highest PoolFactor index must be 0"):
    factor.pool.PoolWiring(*tables, pool_choices_edge_states=choices + np.array([[1, 0]]),
                           pool_indicators_edge_states=indicator).get_inference_arguments()
  two = np.vstack([choices, choices]) + np.array([[0, 0], [1, 0]])
  with pytest.raises(ValueError, match="The PoolWiring must have 1 different PoolFactor indices"):
    factor.pool.PoolWiring(*tables, pool_choices_edge_states=two,
                           pool_indicators_edge_states=indicator).get_inference_arguments()


def _flat(fg):
  """Everything the device plan is built from, in the reference's format."""
  ctx = infer.InfererContext(fg.bp_state)
  out = {"vs": np.asarray(ctx.var_states_for_edge_states), "edge": np.asarray(ctx.edge_indices_for_edge_states),
         "factor": np.asarray(ctx.factor_indices_for_edge_states), "starts": ctx.factor_edge_start}
  for ft, args in ctx.inference_arguments.items():
    for key, value in args.items():
      out[f"{ft.__name__}.{key}"] = np.asarray(value)
  return out


def _assert_same_flat(graphs):
  flats = [_flat(fg) for fg in graphs]
  for other in flats[1:]:
    assert flats[0].keys() == other.keys()
    for key in flats[0]:
      np.testing.assert_array_equal(flats[0][key], other[key], err_msg=key)


def test_wiring_pairwise_group_vs_small_groups_vs_single_factors():
  """tests/fgroup/test_wiring.py:29-101."""
  A = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  B = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  fg1 = fgraph.FactorGraph(variable_groups=[A, B])
  fg1.add_factors(fgroup.PairwiseFactorGroup(variables_for_factors=[[A[i], B[i]] for i in range(10)]))
  assert len(fg1.factor_groups[factor.EnumFactor]) == 1
  fg2 = fgraph.FactorGraph(variable_groups=[A, B])
  for i in range(10):
    fg2.add_factors(fgroup.PairwiseFactorGroup(variables_for_factors=[[A[i], B[i]]]))
  assert len(fg2.factor_groups[factor.EnumFactor]) == 10
  fg3 = fgraph.FactorGraph(variable_groups=[A, B])
  fg3.add_factors([
      factor.EnumFactor(variables=[A[i], B[i]], factor_configs=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]),
                        log_potentials=np.zeros((4,)))
      for i in range(10)
  ])
  assert len(fg3.factor_groups[factor.EnumFactor]) == 10
  assert len(fg1.factors) == len(fg2.factors) == len(fg3.factors)
  _assert_same_flat([fg1, fg2, fg3])


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
def test_wiring_logical_group_vs_small_groups_vs_single_factors(kind):
  """tests/fgroup/test_wiring.py:104-269 (ORFactorGroup, ANDFactorGroup, PoolFactorGroup)."""
  group_cls = {"or": fgroup.ORFactorGroup, "and": fgroup.ANDFactorGroup, "pool": fgroup.PoolFactorGroup}[kind]
  factor_cls = {"or": factor.ORFactor, "and": factor.ANDFactor, "pool": factor.PoolFactor}[kind]
  A = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  B = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  C = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  fg1 = fgraph.FactorGraph(variable_groups=[A, B, C])
  fg1.add_factors(group_cls(variables_for_factors=[[A[i], B[i], C[i]] for i in range(10)]))
  assert len(fg1.factor_groups[factor_cls]) == 1
  fg2 = fgraph.FactorGraph(variable_groups=[A, B, C])
  for i in range(5):
    fg2.add_factors(group_cls(variables_for_factors=[[A[2 * i], B[2 * i], C[2 * i]],
                                                     [A[2 * i + 1], B[2 * i + 1], C[2 * i + 1]]]))
  assert len(fg2.factor_groups[factor_cls]) == 5
  fg3 = fgraph.FactorGraph(variable_groups=[A, B, C])
  for i in range(10):
    fg3.add_factors(factor_cls(variables=[A[i], B[i], C[i]]))
  assert len(fg3.factor_groups[factor_cls]) == 10
  assert len(fg1.factors) == len(fg2.factors) == len(fg3.factors)
  _assert_same_flat([fg1, fg2, fg3])


def test_group_over_variables_with_different_numbers_of_states_is_split_into_uniform_blocks():
  """One EnumFactorGroup may span variables with different numbers of states (its configurations
  only have to be valid for all of them).  The device ABI wants arithmetic message / potential
  offsets inside a block (pgx_enum_block), so the group compiles to one block per run of factors
  with the same state signature - same flat arrays as the reference layout, same factor order."""
  from oracle import bp_oracle

  num_states = np.array([2, 2, 3, 3, 3, 4, 2])
  variables = vgroup.NDVarArray(num_states=num_states, shape=(7,))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = [(0, 1), (0, 6), (2, 3), (3, 4), (5, 2), (1, 6)]  # signatures (2,2) (2,2) (3,3) (3,3) (4,3) (2,2)
  configs = np.array([[0, 0], [1, 1], [0, 1]])
  rng = np.random.RandomState(0)
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[variables[a], variables[b]] for a, b in pairs],
      factor_configs=configs, log_potentials=rng.normal(size=(len(pairs), 3))))
  wiring = fg.bp_state.fg_state.wiring[factor.EnumFactor]
  assert [(b.num_factors, b.first_edge, b.first_config) for b in wiring.blocks] == [
      (2, 0, 0), (2, 4, 6), (1, 8, 12), (1, 10, 15)]
  # the reference-format rows are those of the same factors added one by one
  fg_single = fgraph.FactorGraph(variable_groups=variables)
  lp = np.asarray(fg.bp_state.log_potentials.value).reshape(len(pairs), 3)
  for (a, b), pot in zip(pairs, lp):
    fg_single.add_factors(factor.EnumFactor(variables=[variables[a], variables[b]], factor_configs=configs,
                                            log_potentials=pot))
  _assert_same_flat([fg, fg_single])
  # and the flat-graph form the oracle / the device plan consume gives the same update
  bp = infer.BP(fg.bp_state)
  arrays = bp.init(evidence_updates={variables: rng.gumbel(size=(7, 4))})
  from pgmax_b200 import dist as pdist
  g_ctx = bp_oracle.graph_from_context(bp.context)
  g_flat = bp_oracle.graph_from_flat(pdist.flat_from_state(fg.bp_state.fg_state))
  want, _ = bp_oracle.run_bp(g_ctx, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 5, 0.5, 1.0)
  got, _ = bp_oracle.run_bp(g_flat, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 5, 0.5, 1.0)
  np.testing.assert_array_equal(got, want)
