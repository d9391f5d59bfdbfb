"""Host facade parity with the reference's model-building API (no GPU): the behaviours and
ValueErrors the reference checks in tests/fgraph/test_fgraph.py:31-252, restated against the
NumPy mirror (pgmax_b200.{vgroup,factor,fgroup,fgraph,infer})."""

import re

import numpy as np
import pytest

from pgmax_b200 import factor, fgraph, fgroup, infer, vgroup
from pgmax_b200.infer import bp_state as bpstate


def _one_var_graph(num_configs=10):
  vg = vgroup.VarDict(variable_names=(0,), num_states=15)
  fg = fgraph.FactorGraph(vg)
  group = fgroup.EnumFactorGroup(
      variables_for_factors=[[vg[0]]], factor_configs=np.arange(num_configs)[:, None])
  fg.add_factors(group)
  return vg, fg, group


def test_duplicate_variables_and_duplicate_factors_are_rejected():
  """fgraph.py:96-124 (tests/fgraph/test_fgraph.py:31-75)."""
  vg = vgroup.VarDict(variable_names=(0,), num_states=15)
  fg = fgraph.FactorGraph(vg)
  twice = factor.EnumFactor(
      variables=[vg[0], vg[0]], factor_configs=np.array([[i, i] for i in range(15)]),
      log_potentials=np.zeros(15))
  with pytest.raises(ValueError, match=re.escape(
      f"A Factor of type {factor.EnumFactor} involving variables"
      f" {[(vg.__hash__(), 15), (vg.__hash__(), 15)]} contains variables duplicates.")):
    fg.add_factors(twice)
  fg.add_factors(factor.EnumFactor(
      variables=[vg[0]], factor_configs=np.arange(15)[:, None], log_potentials=np.zeros(15)))
  again = fgroup.EnumFactorGroup(
      variables_for_factors=[[vg[0]]], factor_configs=np.arange(15)[:, None], log_potentials=np.zeros(15))
  with pytest.raises(ValueError, match=re.escape(
      f"A Factor of type {factor.EnumFactor} involving variables"
      f" {frozenset([(vg.__hash__(), 15)])} already exists.")):
    fg.add_factors(again)


def test_bp_state_parts_must_share_the_fg_state():
  """bp_state.py:426-432 (tests/fgraph/test_fgraph.py:78-101)."""
  vg = vgroup.VarDict(variable_names=(0,), num_states=15)
  single = factor.EnumFactor(
      variables=[vg[0]], factor_configs=np.arange(15)[:, None], log_potentials=np.zeros(15))
  fg0, fg1 = fgraph.FactorGraph(vg), fgraph.FactorGraph(vg)
  fg0.add_factors(single)
  fg1.add_factors(single)
  with pytest.raises(ValueError, match=(
      "log_potentials, ftov_msgs and evidence should be derived from the same fg_state")):
    infer.BPState(log_potentials=fg0.bp_state.log_potentials, ftov_msgs=fg1.bp_state.ftov_msgs,
                  evidence=fg1.bp_state.evidence)


def test_log_potentials_errors_and_lookup():
  """bp_state.py:87-126 (tests/fgraph/test_fgraph.py:104-141)."""
  vg, fg, group = _one_var_graph()
  with pytest.raises(ValueError, match=re.escape("Expected log potentials shape (10,) for factor group.")):
    fg.bp_state.log_potentials[group] = np.zeros((1, 15))
  with pytest.raises(ValueError, match=re.escape("Invalid FactorGroup for log potentials updates.")):
    other = fgroup.EnumFactorGroup(
        variables_for_factors=[[vg[0]]], factor_configs=np.arange(10)[:, None])
    fg.bp_state.log_potentials[other] = np.zeros((1, 15))
  with pytest.raises(ValueError, match=re.escape("Invalid FactorGroup queried to access log potentials.")):
    _ = fg.bp_state.log_potentials[vg[0]]
  with pytest.raises(ValueError, match=re.escape("Expected log potentials shape (10,). Got (15,)")):
    infer.LogPotentials(fg_state=fg.fg_state, value=np.zeros(15))
  lp = infer.LogPotentials(fg_state=fg.fg_state, value=np.zeros((10,)))
  assert np.all(lp[group] == np.zeros((10,)))


def test_ftov_msgs_errors():
  """bp_state.py:198-262 (tests/fgraph/test_fgraph.py:144-197)."""
  vg, fg, group = _one_var_graph()
  with pytest.raises(ValueError, match=re.escape("Provided variable or factor type is not in the FactorGraph")):
    fg.bp_state.ftov_msgs[0] = np.ones(10)
  with pytest.raises(ValueError, match=re.escape(
      f"Expected ftov_msgs shape (15,) for variable ({vg.__hash__()}, 15). Got incompatible shape (10,).")):
    fg.bp_state.ftov_msgs[vg[0]] = np.ones(10)
  with pytest.raises(ValueError, match=re.escape("Expected messages shape (15,). Got (10,)")):
    infer.FToVMessages(fg_state=fg.fg_state, value=np.zeros(10))
  msgs = infer.FToVMessages(fg_state=fg.fg_state, value=np.zeros(15))
  with pytest.raises(TypeError, match=re.escape("'FToVMessages' object is not subscriptable")):
    _ = msgs[(10,)]
  with pytest.raises(ValueError, match=re.escape(
      f"Expected ftov_msgs shape (15,) for factor type {factor.EnumFactor}. Got incompatible shape (10,).")):
    bpstate.update_ftov_msgs(msgs.value, {factor.EnumFactor: np.zeros(10)}, fg.fg_state)


def test_evidence_errors():
  """bp_state.py:318-367 (tests/fgraph/test_fgraph.py:200-252)."""
  vg = vgroup.VarDict(variable_names=(0, 1), num_states=15)
  fg = fgraph.FactorGraph(vg)
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[vg[0], vg[1]]],
      factor_configs=(np.arange(10)[:, None] + np.zeros((1, 2))).astype(int)))
  evidence = infer.Evidence(fg_state=fg.fg_state, value=np.zeros((30,)))
  assert np.all(evidence.value == np.zeros((30,)))
  with pytest.raises(ValueError, match=re.escape("Expected evidence shape (30,). Got (20,).")):
    infer.Evidence(fg_state=fg.fg_state, value=np.zeros(20))
  with pytest.raises(ValueError, match=re.escape(
      f"Expected evidence shape (15,) for variable {vg[0]}. Got incompatible shape (10,).")):
    bpstate.update_evidence(evidence.value, {vg[0]: np.zeros(10)}, fg.fg_state)
  stranger = vgroup.VarDict(variable_names=(0,), num_states=15)
  with pytest.raises(ValueError, match=re.escape(
      "Got evidence for a variable or a VarGroup not in the FactorGraph!")):
    bpstate.update_evidence(evidence.value, {stranger[0]: np.zeros(15)}, fg.fg_state)


def test_update_and_to_bp_state_without_a_device():
  """bp.update / to_bp_state are host-side (inferer.py:120-209; tests/fgraph/test_fgraph.py:255-287
  up to the run): potentials and evidence land at the right offsets, to_bp_state round-trips."""
  vg, fg, group = _one_var_graph()
  rng = np.random.default_rng(0)
  ev = {var: rng.gumbel(size=(var[1],)) for var in vg.variables}
  bp = infer.build_inferer(fg.bp_state, backend="bp")
  arrays = bp.update()
  arrays = bp.update(bp_arrays=arrays, log_potentials_updates={group: np.ones(10)}, evidence_updates=ev)
  np.testing.assert_allclose(arrays.log_potentials, np.ones(10))
  np.testing.assert_allclose(arrays.evidence, ev[vg[0]].astype(np.float32), rtol=1e-6)
  state = bp.to_bp_state(arrays)
  assert state.fg_state == fg.fg_state
  np.testing.assert_allclose(state.evidence.value, arrays.evidence)


def test_ragged_num_states_flat_layout():
  """VarDict and NDVarArray with different numbers of states per variable
  (tests/fgraph/test_fgraph.py:290-333, host part): evidence updates by variable and by group
  fill every var-state exactly once; beliefs unflatten to padded arrays."""
  num_states = np.array([2, 3, 4])
  vdict = vgroup.VarDict(variable_names=("a", "b", "c"), num_states=num_states)
  varray = vgroup.NDVarArray(shape=(3,), num_states=num_states)
  fg = fgraph.FactorGraph([vdict, varray])
  for name, idx, ns in zip(["a", "b", "c"], [0, 1, 2], num_states):
    fg.add_factors(factor.EnumFactor(
        variables=[vdict[name], varray[idx]], factor_configs=np.array([[s, s] for s in range(ns)]),
        log_potentials=np.zeros(ns)))
  bp = infer.build_inferer(fg.bp_state, backend="bp")
  rng = np.random.default_rng(1)
  arrays = bp.init(evidence_updates={var: rng.gumbel(size=(var[1],)) for var in vdict.variables})
  arrays = bp.update(bp_arrays=arrays, evidence_updates={varray: rng.gumbel(size=(3, num_states.max()))})
  assert arrays.evidence.shape == (2 * int(num_states.sum()),)
  assert np.all(arrays.evidence != 0)
  beliefs = infer.inferer.unflatten_beliefs(arrays.evidence, fg.fg_state.variable_groups)
  assert beliefs[varray].shape == (3, 4) and np.isneginf(beliefs[varray][0, 2:]).all()
  decoded = infer.decode_map_states(beliefs)
  assert set(decoded[vdict].keys()) == {"a", "b", "c"}


def test_single_and_empty_factor_groups():
  """fgroup.py:60-140 (tests/fgroup/test_fgroup.py:24-49)."""
  with pytest.raises(ValueError, match="Cannot create a FactorGroup with no Factor."):
    fgroup.ORFactorGroup(variables_for_factors=[])
  A = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  B = vgroup.NDVarArray(num_states=2, shape=(10,))  # pylint: disable=invalid-name
  first, second = (A[0], B[0]), (A[1], B[1])
  or0 = fgroup.ORFactorGroup(variables_for_factors=[first])
  with pytest.raises(ValueError, match="SingleFactorGroup should only contain one factor. Got 2"):
    fgroup.SingleFactorGroup(variables_for_factors=[first, second], single_factor=or0)
  or1 = fgroup.ORFactorGroup(variables_for_factors=[second])
  assert (or0 < or1) in (True, False)  # factor groups are orderable


def test_enum_factor_group_shapes_and_lookup():
  """fgroup/enum.py:60-207 (tests/fgroup/test_fgroup.py:52-135)."""
  vg = vgroup.NDVarArray(shape=(2, 2), num_states=3)
  triples = [[vg[0, 0], vg[0, 1], vg[1, 1]], [vg[0, 1], vg[1, 0], vg[1, 1]]]
  with pytest.raises(ValueError, match=re.escape("Expected log potentials shape: (1,) or (2, 1). Got (3, 2)")):
    fgroup.EnumFactorGroup(variables_for_factors=triples, factor_configs=np.zeros((1, 3), dtype=int),
                           log_potentials=np.zeros((3, 2)))
  with pytest.raises(ValueError, match=re.escape("Potentials should be floats")):
    fgroup.EnumFactorGroup(variables_for_factors=triples, factor_configs=np.zeros((1, 3), dtype=int),
                           log_potentials=np.zeros((2, 1), dtype=int))
  group = fgroup.EnumFactorGroup(variables_for_factors=triples, factor_configs=np.zeros((1, 3), dtype=int))
  missing = [vg[0, 0], vg[1, 1]]
  with pytest.raises(ValueError, match=re.escape(
      f"The queried factor connected to the set of variables {frozenset(missing)} is not present in the"
      " factor group.")):
    _ = group[missing]
  assert group[[vg[0, 1], vg[1, 0], vg[1, 1]]] == group.factors[1]
  with pytest.raises(ValueError, match=re.escape("data should be of shape (2, 1) or (2, 9) or (1,). Got (4, 5).")):
    group.flatten(np.zeros((4, 5)))
  assert np.all(group.flatten(np.ones(1)) == np.ones(2))
  assert np.all(group.flatten(np.ones((2, 9))) == np.ones(18))
  with pytest.raises(ValueError, match=re.escape("Can only unflatten 1D array. Got a 3D array.")):
    group.unflatten(np.ones((1, 2, 3)))
  with pytest.raises(ValueError, match=re.escape(
      "flat_data should be compatible with shape (2, 1) or (2, 9). Got (30,)")):
    group.unflatten(np.zeros(30))
  assert np.all(group.unflatten(np.arange(2)) == np.array([[0], [1]]))
  assert np.all(group.unflatten(np.ones(18)) == np.ones((2, 9)))


def test_pairwise_factor_group_shapes():
  """fgroup/enum.py:210-435 (tests/fgroup/test_fgroup.py:138-230)."""
  vg = vgroup.NDVarArray(shape=(2, 2), num_states=3)
  with pytest.raises(ValueError, match=re.escape("log_potential_matrix should be either a 2D array")):
    fgroup.PairwiseFactorGroup([[vg[0, 0], vg[1, 1]]], np.zeros((1,), dtype=float))
  with pytest.raises(ValueError, match=re.escape("Potential matrix should be floats")):
    fgroup.PairwiseFactorGroup([[vg[0, 0], vg[1, 1]]], np.zeros((3, 3), dtype=int))
  with pytest.raises(ValueError, match=re.escape(
      "Expected log_potential_matrix for 1 factors. Got log_potential_matrix for 2 factors.")):
    fgroup.PairwiseFactorGroup([[vg[0, 0], vg[1, 1]]], np.zeros((2, 3, 3), dtype=float))
  with pytest.raises(ValueError, match=re.escape(
      "All pairwise factors should connect to exactly 2 variables. Got a factor connecting to 3 variables")):
    fgroup.PairwiseFactorGroup([[vg[0, 0], vg[1, 1], vg[0, 1]]], np.zeros((3, 3), dtype=float))
  pair = [vg[0, 0], vg[1, 1]]
  with pytest.raises(ValueError, match=re.escape(f"The specified pairwise factor {pair}")):
    fgroup.PairwiseFactorGroup([pair], np.zeros((4, 4), dtype=float))
  group = fgroup.PairwiseFactorGroup([[vg[0, 0], vg[1, 1]], [vg[1, 0], vg[0, 1]]])
  with pytest.raises(ValueError, match=re.escape(
      "data should be of shape (2, 3, 3) or (2, 6) or (3, 3). Got (4, 4).")):
    group.flatten(np.zeros((4, 4)))
  assert np.all(group.flatten(np.zeros((3, 3))) == np.zeros(2 * 3 * 3))
  assert np.all(group.flatten(np.zeros((2, 6))) == np.zeros(12))
  with pytest.raises(ValueError, match="Can only unflatten 1D array. Got a 2D array"):
    group.unflatten(np.zeros((10, 20)))
  assert np.all(group.unflatten(np.zeros(2 * 3 * 3)) == np.zeros((2, 3, 3)))
  assert np.all(group.unflatten(np.zeros(2 * 6)) == np.zeros((2, 6)))
  with pytest.raises(ValueError, match=re.escape(
      "flat_data should be compatible with shape (2, 3, 3) or (2, 6). Got (10,).")):
    group.unflatten(np.zeros(10))


def test_var_dict_checks():
  """vgroup/vdict.py (tests/vgroup/test_vgroup.py:24-88)."""
  with pytest.raises(ValueError, match=re.escape("Expected num_states shape (3,). Got (4,).")):
    vgroup.VarDict(variable_names=(0, 1, 2), num_states=np.full((4,), 2))
  with pytest.raises(ValueError, match=re.escape("num_states should be an integer or a NumPy array of dtype int")):
    vgroup.VarDict(variable_names=(0, 1, 2), num_states=np.full((3,), 2, dtype=np.float32))
  vd = vgroup.VarDict(variable_names=(0, 1, 2), num_states=15)
  with pytest.raises(ValueError, match="data is referring to a non-existent variable 3"):
    vd.flatten({3: np.zeros(10)})
  with pytest.raises(ValueError, match=re.escape(
      "Variable 2 expects a data array of shape (15,) or (1,). Got (10,).")):
    vd.flatten({2: np.zeros(10)})
  with pytest.raises(ValueError, match="Can only unflatten 1D array. Got a 2D array."):
    vd.unflatten(np.zeros((10, 20)), True)
  per_var = vd.unflatten(np.zeros(3), False)
  assert set(per_var.keys()) == {0, 1, 2} and all(np.all(v == 0) for v in per_var.values())
  with pytest.raises(ValueError, match=re.escape(
      "flat_data should be shape (num_variable_states(=45),). Got (100,)")):
    vd.unflatten(np.zeros(100), True)
  with pytest.raises(ValueError, match=re.escape("flat_data should be shape (num_variables(=3),). Got (100,)")):
    vd.unflatten(np.zeros(100), False)


def test_nd_var_array_checks():
  """vgroup/varray.py (tests/vgroup/test_vgroup.py:91-187)."""
  max_size = int(vgroup.vgroup.MAX_SIZE)
  with pytest.raises(ValueError, match=re.escape(
      f"Currently only support NDVarArray of size smaller than {max_size}. Got {max_size + 1}")):
    vgroup.NDVarArray(shape=(max_size + 1,), num_states=2)
  with pytest.raises(ValueError, match=re.escape("Expected num_states shape (2, 2). Got (2, 3).")):
    vgroup.NDVarArray(shape=(2, 2), num_states=np.full((2, 3), 2))
  with pytest.raises(ValueError, match=re.escape("num_states should be an integer or a NumPy array of dtype int")):
    vgroup.NDVarArray(shape=(2, 2), num_states=np.full((2, 3), 2, dtype=np.float32))
  grid = vgroup.NDVarArray(shape=(5, 5), num_states=2)
  assert len(grid[:3, :3]) == 9
  ragged = vgroup.NDVarArray(shape=(2, 2), num_states=np.array([[1, 2], [3, 4]]))
  assert repr(grid) and repr(ragged)
  assert (grid < ragged) in (True, False)
  with pytest.raises(ValueError, match=re.escape("data should be of shape (2, 2) or (2, 2, 4). Got (3, 3).")):
    ragged.flatten(np.zeros((3, 3)))
  assert np.all(ragged.flatten(np.array([[1, 2], [3, 4]])) == np.array([1, 2, 3, 4]))
  assert np.all(ragged.flatten(np.zeros((2, 2, 4))) == np.zeros((10,)))
  # deliberate extension: a 2-D input is (batch, flat) - the leading batch axis that replaces
  # jax.vmap; anything beyond that is rejected with the reference's message
  assert ragged.unflatten(np.zeros((3, 10)), True).shape == (3, 2, 2, 4)
  with pytest.raises(ValueError, match="Can only unflatten 1D array. Got a 3D array."):
    ragged.unflatten(np.zeros((2, 3, 4)), True)
  with pytest.raises(ValueError, match=re.escape("flat_data size should be equal to 10. Got size 12.")):
    ragged.unflatten(np.zeros((12,)), True)
  with pytest.raises(ValueError, match=re.escape("flat_data size should be equal to 4. Got size 12.")):
    ragged.unflatten(np.zeros((12,)), False)
  assert np.all(ragged.unflatten(np.zeros(4), False) == np.zeros((2, 2)))
  padded = ragged.unflatten(np.zeros(10), True)
  assert padded.shape == (2, 2, 4)
  assert np.all(padded[0, 0, :1] == 0) and np.all(padded[0, 1, :2] == 0)
  assert np.all(padded[1, 0, :3] == 0) and np.all(padded[1, 1] == 0)


def test_nd_var_array_repr():
  """tests/vgroup/test_vgroup.py:190-201."""
  assert repr(vgroup.NDVarArray(shape=(0,), num_states=np.zeros((0,), int))).startswith(
      "NDVarArray(shape=(0,), num_states=[]")
  assert repr(vgroup.NDVarArray(shape=(2,), num_states=np.array([2, 2]))).startswith(
      "NDVarArray(shape=(2,), num_states=2")
  assert repr(vgroup.NDVarArray(shape=(2,), num_states=np.array([2, 3]))).startswith(
      "NDVarArray(shape=(2,), min_num_states=2, max_num_states=3")


def test_sdlp_facade_without_a_device():
  """build_inferer dispatch (pgmax/infer/__init__.py:34-40), the SDLP argument checks
  (pgmax/infer/dual_lp.py:266-277; tests/lp/test_dual_lp.py:70-90) and the host-side pieces:
  factor membership of the edges and the per-iteration fp32 scalars."""
  from pgmax_b200.infer import dual_lp
  import models

  fg, variables = models.sdlp_ising_model()
  sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
  assert isinstance(sdlp, infer.SmoothDualLP)
  with pytest.raises(NotImplementedError, match="Inferer other is not supported."):
    infer.build_inferer(fg.bp_state, backend="other")
  arrays = sdlp.init()
  with pytest.raises(ValueError, match="has to be between 0.0 and 1.0"):
    sdlp.run(arrays, logsumexp_temp=1.01, num_iters=1)
  with pytest.raises(ValueError, match="the learning rate must be smaller than the log sum-exp temperature"):
    sdlp.run(arrays, logsumexp_temp=0.01, num_iters=1, lr=0.1)
  ctx = sdlp.context
  starts = ctx.factor_edge_start
  assert starts.shape == (ctx.num_factors + 1,) and starts[0] == 0 and starts[-1] == ctx.num_edges
  # the same membership as the reference's per-edge-state column (inferer.py:76-98)
  edge_of_es, factor_of_es = ctx.edge_indices_for_edge_states, ctx.factor_indices_for_edge_states
  np.testing.assert_array_equal(np.searchsorted(starts, edge_of_es, side="right") - 1, factor_of_es)
  steps, momenta = dual_lp.step_schedule(4, 0.01, 0.0)
  np.testing.assert_allclose(steps, 0.01 / np.sqrt(np.arange(1, 5)), rtol=1e-6)
  np.testing.assert_allclose(momenta, np.arange(1, 5) / np.arange(4, 8), rtol=1e-6)
  steps, _ = dual_lp.step_schedule(4, 0.25, 0.5)
  assert np.all(steps == np.float32(0.25))


def test_heretic_model_construction_and_marginals():
  """tests/test_pgmax.py:424-475 ("heretic" model): 9 PairwiseFactorGroups between 17-state hidden
  and 3-state pixel variables, evidence set per group and per variable, 7056 factors; one
  sum-product iteration (here on the oracle - the device run is in the GPU suite) and
  get_marginals sums to one."""
  from oracle import bp_oracle
  from pgmax_b200.infer import inferer

  im_size = (30, 30)
  pixel_vars = vgroup.NDVarArray(shape=im_size, num_states=3)
  hidden_vars = vgroup.NDVarArray(shape=(im_size[0] - 2, im_size[1] - 2), num_states=17)
  fg = fgraph.FactorGraph([pixel_vars, hidden_vars])
  w_pot = np.random.RandomState(0).normal(size=(17, 3, 3, 3))
  for k_row in range(3):
    for k_col in range(3):
      fg.add_factors(fgroup.PairwiseFactorGroup(
          variables_for_factors=[[hidden_vars[r, c], pixel_vars[r + k_row, c + k_col]]
                                 for r in range(28) for c in range(28)],
          log_potential_matrix=w_pot[:, :, k_row, k_col]))
  bp_state = fg.bp_state
  bp_state.evidence[pixel_vars] = np.zeros((30, 30, 3))
  bp_state.evidence[pixel_vars[0, 0]] = np.array([0.0, 0.0, 0.0])
  assert bp_state.evidence[pixel_vars[0, 0]].shape == (3,)
  assert bp_state.evidence[hidden_vars[0, 0]].shape == (17,)
  assert isinstance(bp_state.evidence.value, np.ndarray)
  assert len(sum(fg.factors.values(), ())) == 7056
  bp = infer.build_inferer(bp_state, backend="bp")
  arrays = bp.init()
  graph = bp_oracle.graph_from_context(bp.context)
  msgs, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1, 0.5, 1.0)
  beliefs = inferer.unflatten_beliefs(bp_oracle.flat_beliefs(graph, msgs, arrays.evidence),
                                      bp_state.fg_state.variable_groups)
  marginals = infer.get_marginals(beliefs)
  np.testing.assert_allclose(marginals[pixel_vars].sum(axis=-1), 1.0, atol=1e-6)
  np.testing.assert_allclose(marginals[hidden_vars].sum(axis=-1), 1.0, atol=1e-6)


def test_vardict_unflatten_batch_keeps_the_batch_axis():
  """The batch axis that replaces jax.vmap reaches a VarDict as [B, n] arrays (through
  get_beliefs / unflatten_states -> VarDict.unflatten_batch); VarDict.unflatten itself keeps the
  reference's 1-D contract (pgmax/vgroup/vdict.py:129-189, test_var_dict_checks above)."""
  vd = vgroup.VarDict(variable_names=("a", "b"), num_states=np.array([2, 3]))
  flat = np.arange(10, dtype=np.float32).reshape(2, 5)
  out = vd.unflatten_batch(flat, True)
  np.testing.assert_array_equal(out["a"], flat[:, :2])
  np.testing.assert_array_equal(out["b"], flat[:, 2:])
  per_var = vd.unflatten_batch(np.array([[0, 2], [1, 1]]), False)
  np.testing.assert_array_equal(per_var["a"], [0, 1])
  np.testing.assert_array_equal(per_var["b"], [2, 1])
  beliefs = infer.inferer.unflatten_beliefs(flat, [vd])
  np.testing.assert_array_equal(beliefs[vd]["b"], flat[:, 2:])
  with pytest.raises(ValueError, match="batch, flat"):
    vd.unflatten_batch(np.zeros((2, 2, 5)), True)
  with pytest.raises(ValueError, match="flat_data should be shape"):
    vd.unflatten_batch(np.zeros((2, 4)), True)


def test_bp_arrays_do_not_freeze_the_callers_buffers():
  """BPArrays are immutable like the reference's (pgmax/infer/bp_state.py:45-48) but hold
  read-only VIEWS: the arrays the user passed in stay writable."""
  from pgmax_b200.infer.bp_state import BPArrays
  lp, msgs, ev = (np.zeros(4, np.float32) for _ in range(3))
  arrays = BPArrays(log_potentials=lp, ftov_msgs=msgs, evidence=ev)
  assert not arrays.evidence.flags.writeable
  with pytest.raises(ValueError):
    arrays.evidence[0] = 1.0
  ev[0] = 2.0  # still the caller's buffer
  assert lp.flags.writeable and msgs.flags.writeable and ev.flags.writeable
  assert arrays.evidence[0] == 2.0
