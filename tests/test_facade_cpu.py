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
