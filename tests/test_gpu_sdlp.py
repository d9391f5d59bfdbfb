"""GPU parity of the smooth dual LP-MAP solver (pgx_sdlp_* behind pgmax_b200.infer.SDLP)
against the CPU oracle (oracle/sdlp_oracle.py) and the reference's own properties
(tests/lp/test_dual_lp.py).  Every call goes through the C ABI.

Tolerances: T = 0 quantities (max / arg-max based, no transcendental) at 1e-6 absolute;
T > 0 at 1e-5 absolute for T >= 0.1 (north-star sum-product tolerance); at T = 1e-3 the
softmax argument is divided by T, so one fp32 ulp of a message moves the gradient by up to
~1e-3 * |g|: gradient compared at 2e-3 there, objective and logsumexps stay at 1e-5 relative.
"""

import os

import numpy as np
import pytest

import models
from oracle import bp_oracle
from oracle import sdlp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup

pytestmark = pytest.mark.gpu
RTOL = 5e-3  # tests/lp/test_dual_lp.py:27
# LP optima of the reference's test models from an independent solver (tests/golden/make_sdlp_golden.py)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdlp_lp.npz"))


def _check_lp(name, sdlp, decoded, upper, lower):
  """tests/lp/test_dual_lp.py:123,221,304,395 (bound == LP optimum, rtol 5e-3) against the committed
  HiGHS optimum; the relaxations are tight, so the device's decoding is the LP's integral solution."""
  objval = float(GOLD[f"{name}_objval"])
  assert np.isclose(objval, upper, rtol=RTOL), (objval, upper)
  assert np.isclose(objval, lower, rtol=RTOL), (objval, lower)
  info = sdlp.context.bp_state.fg_state
  states = np.concatenate([np.asarray(decoded[vg]).reshape(-1) for vg in info.variable_groups])
  bounds = np.concatenate([[0], np.cumsum(np.concatenate([vg.num_states.reshape(-1) for vg in info.variable_groups]))])
  lp_states = np.array([int(np.argmax(GOLD[f"{name}_solution"][bounds[v] : bounds[v + 1]]))
                        for v in range(len(bounds) - 1)])
  np.testing.assert_array_equal(states, lp_states)


def _tols(temp):
  if temp == 0.0:
    return dict(vals=1e-6, grad=1e-6)
  if temp < 0.1:
    return dict(vals=1e-4, grad=2e-3)
  return dict(vals=1e-5, grad=1e-5)


def _pick(a, b):
  a = np.asarray(a)
  return a[b] if a.ndim == 2 else a


def _check_eval(sdlp, arrays, temp, batch=None):
  graph = bp_oracle.graph_from_context(sdlp.context)
  got = sdlp.objval_and_grad(arrays, temp, want=("grad", "bp_updates", "edge_vals"))
  tol = _tols(temp)
  for b in range(batch or 1):
    objval, grad, updates, edge_vals = sdlp_oracle.smooth_dual_objval_and_grad(
        graph, _pick(arrays.ftov_msgs, b), _pick(arrays.log_potentials, b), _pick(arrays.evidence, b), temp)
    sel = (lambda x: np.asarray(x)[b]) if batch else (lambda x: np.asarray(x))
    finite = np.isfinite(updates)
    np.testing.assert_array_equal(np.isfinite(sel(got["bp_updates"])), finite)
    np.testing.assert_allclose(sel(got["bp_updates"])[finite], updates[finite], atol=tol["vals"], rtol=1e-5)
    np.testing.assert_allclose(sel(got["edge_vals"]), edge_vals, atol=tol["vals"], rtol=1e-5)
    np.testing.assert_allclose(sel(got["grad"]), grad, atol=tol["grad"])
    np.testing.assert_allclose(sel(got["objval"]), objval, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("temp", [0.0, 1e-3, 0.5, 1.0])
@pytest.mark.parametrize("batch", [None, 3, 40])
def test_objval_and_grad_pairwise_enum(temp, batch):
  fg, variables = models.sdlp_ising_model(seed=2, scale=1.0)
  sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
  rng = np.random.RandomState(5)
  shape = (4, 4, 3) if batch is None else (batch, 4, 4, 3)
  arrays = sdlp.init(evidence_updates={variables: rng.gumbel(size=shape)})
  msgs = rng.normal(size=((batch,) if batch else ()) + (arrays.ftov_msgs.shape[-1],)).astype(np.float32)
  arrays = infer.BPArrays(log_potentials=arrays.log_potentials, ftov_msgs=msgs, evidence=arrays.evidence)
  _check_eval(sdlp, arrays, temp, batch)


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
@pytest.mark.parametrize("temp", [0.0, 0.3, 1.0])
def test_objval_and_grad_logical_and_enum(kind, temp):
  for seed in (1, 2):
    data = models.logical_pair(kind, seed)
    entry = data["graphs"][0]
    sdlp = infer.SDLP(entry[0].bp_state)
    arrays = models.init_logical(sdlp, entry, data)
    _check_eval(sdlp, arrays, temp)


@pytest.mark.parametrize("temp", [0.0, 1.0])
def test_objval_and_grad_large_enum_factors(temp):
  """Factors with more than 64 edge-states take the CTA-per-factor kernel (k_enum_big<raw>);
  ragged numbers of states, a subset of the configurations valid."""
  rng = np.random.RandomState(0)
  variables = vgroup.NDVarArray(num_states=np.array([40, 50, 45]), shape=(3,))
  fg = fgraph.FactorGraph(variable_groups=variables)
  for a, b in ((0, 1), (1, 2), (0, 2)):
    na, nb = int(variables.num_states[a]), int(variables.num_states[b])
    configs = np.array([(i, j) for i in range(na) for j in range(nb) if (i * 7 + j * 3) % 5 != 0])
    fg.add_factors(fgroup.EnumFactorGroup(
        variables_for_factors=[[variables[a], variables[b]]], factor_configs=configs,
        log_potentials=rng.normal(size=(configs.shape[0],))))
  sdlp = infer.SDLP(fg.bp_state)
  arrays = sdlp.init(evidence_updates={variables: rng.gumbel(size=(3, 50))})
  msgs = rng.normal(size=arrays.ftov_msgs.shape).astype(np.float32)
  arrays = infer.BPArrays(log_potentials=arrays.log_potentials, ftov_msgs=msgs, evidence=arrays.evidence)
  _check_eval(sdlp, arrays, temp)


# The default learning rate (lr = T) sits above 1 / L of the summed variable + factor terms on
# these graphs: the oracle run with its input perturbed by 1e-6 is 0.6 away after 60 iterations
# (and 4e-6 after 5), so multi-iteration parity uses lr = 0.1 (perturbation stays at 2e-6) and the
# default lr is compared over a few iterations only.
@pytest.mark.parametrize("temp,num_iters,lr", [(0.0, 40, None), (0.5, 60, 0.1), (0.5, 6, None), (1.0, 6, None)])
@pytest.mark.parametrize("batch", [None, 5])
def test_run_matches_oracle(temp, num_iters, lr, batch):
  fg, variables = models.sdlp_ising_model(seed=4, scale=1.0)
  sdlp = infer.SDLP(fg.bp_state)
  rng = np.random.RandomState(9)
  shape = (4, 4, 3) if batch is None else (batch, 4, 4, 3)
  arrays = sdlp.init(evidence_updates={variables: rng.gumbel(size=shape)})
  out, objvals = sdlp.run_with_objvals(arrays, logsumexp_temp=temp, num_iters=num_iters, lr=lr)
  graph = bp_oracle.graph_from_context(sdlp.context)
  for b in range(batch or 1):
    want, want_obj = sdlp_oracle.run_with_objvals(
        graph, arrays.log_potentials, arrays.ftov_msgs, _pick(arrays.evidence, b), temp, num_iters, lr)
    got = np.asarray(out.ftov_msgs)[b] if batch else np.asarray(out.ftov_msgs)
    got_obj = np.asarray(objvals)[b] if batch else np.asarray(objvals)
    np.testing.assert_allclose(got, want, atol=1e-6 if temp == 0.0 else 2e-5)
    np.testing.assert_allclose(got_obj, want_obj, rtol=1e-5, atol=1e-4)


def test_run_with_or_factors_matches_oracle():
  fg, top, bottom, evidence = models.sdlp_line_model(seed=1)
  sdlp = infer.SDLP(fg.bp_state)
  arrays = sdlp.init(evidence_updates=evidence)
  graph = bp_oracle.graph_from_context(sdlp.context)
  for temp, lr, num_iters in ((0.0, None, 30), (0.5, 0.02, 30), (0.5, 0.1, 10), (0.5, None, 4)):
    out, objvals = sdlp.run_with_objvals(arrays, logsumexp_temp=temp, num_iters=num_iters, lr=lr)
    want, want_obj = sdlp_oracle.run_with_objvals(
        graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, temp, num_iters, lr)
    np.testing.assert_allclose(out.ftov_msgs, want, atol=2e-3, rtol=1e-5)  # evidence of 1e4: ulp 1e-3
    np.testing.assert_allclose(objvals, want_obj, rtol=1e-5)


@pytest.mark.parametrize("seed,temp", [(0, 1e-3), (1, 0.0)])
def test_dual_bounds_meet_on_tight_ising(seed, temp):
  """tests/lp/test_dual_lp.py:30-135 without cvxpy: upper and lower bound of the primal meet."""
  fg, variables = models.sdlp_ising_model(seed=seed)
  sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
  rng = np.random.RandomState(seed)
  arrays = sdlp.init(evidence_updates={variables: rng.gumbel(size=(4, 4, 3))})
  arrays = sdlp.run(arrays, logsumexp_temp=temp, lr=None, num_iters=5000)
  decoded, _ = sdlp.decode_primal_unaries(arrays)
  upper = sdlp.get_primal_upper_bound(arrays)
  lower = sdlp.get_map_lower_bound(arrays, decoded)
  assert np.isclose(lower, upper, rtol=RTOL)
  assert np.isclose(lower, sdlp.get_map_lower_bound(arrays, decoded, debug_mode=True))
  _check_lp(f"ising_{seed}", sdlp, decoded, upper, lower)


def test_line_sparsification():
  """tests/lp/test_dual_lp.py:139-235."""
  for seed in range(3):
    fg, top, bottom, evidence = models.sdlp_line_model(seed=seed)
    sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
    arrays = sdlp.init(evidence_updates=evidence)
    arrays = sdlp.run(arrays, logsumexp_temp=1e-3, lr=None, num_iters=5000)
    decoded, _ = sdlp.decode_primal_unaries(arrays)
    assert decoded[top].sum() == (20 + 3) // 3
    upper = sdlp.get_primal_upper_bound(arrays)
    lower = sdlp.get_map_lower_bound(arrays, decoded)
    assert np.isclose(lower, upper, rtol=RTOL)
    _check_lp(f"line_{seed}", sdlp, decoded, upper, lower)


@pytest.mark.parametrize("seed", [0, 1])
def test_and_and_pool_models_reach_the_lp_optimum(seed):
  """tests/lp/test_dual_lp.py:231-402 on the device: ANDFactors (rows of ones) and the PoolFactor
  hierarchy, bounds and decodings against the LP fixture."""
  fg, matrix, all_ones, evidence, truth = models.sdlp_and_model(seed=seed)
  sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
  arrays = sdlp.run(sdlp.init(evidence_updates=evidence), logsumexp_temp=1e-3, lr=None, num_iters=5000)
  decoded, _ = sdlp.decode_primal_unaries(arrays)
  np.testing.assert_array_equal(np.asarray(decoded[all_ones]).reshape(-1), truth)
  _check_lp(f"and_{seed}", sdlp, decoded, sdlp.get_primal_upper_bound(arrays), sdlp.get_map_lower_bound(arrays, decoded))
  fg, variables = models.sdlp_pool_model()
  updates = np.random.RandomState(seed).gumbel(size=(variables.shape[0], 2))
  updates[0, 1] = 1_000
  sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
  arrays = sdlp.run(sdlp.init(evidence_updates={variables: updates}), logsumexp_temp=1e-3, lr=None, num_iters=5000)
  decoded, _ = sdlp.decode_primal_unaries(arrays)
  assert int(np.asarray(decoded[variables]).sum()) == 4
  _check_lp(f"pool_{seed}", sdlp, decoded, sdlp.get_primal_upper_bound(arrays), sdlp.get_map_lower_bound(arrays, decoded))


def test_zero_iterations_and_shared_messages():
  fg, variables = models.sdlp_ising_model(seed=0)
  sdlp = infer.SDLP(fg.bp_state)
  rng = np.random.RandomState(0)
  arrays = sdlp.init(evidence_updates={variables: rng.gumbel(size=(3, 4, 4, 3))})
  out, objvals = sdlp.run_with_objvals(arrays, logsumexp_temp=0.5, num_iters=0)
  assert np.asarray(objvals).shape == (3, 0)
  np.testing.assert_array_equal(np.asarray(out.ftov_msgs), np.zeros((3, arrays.ftov_msgs.shape[-1]), np.float32))


@pytest.mark.parametrize("temperature", [1e-3, 0.01])
def test_low_temperature_updates_and_per_factor_consistency(temperature):
  """tests/lp/test_bp_for_lp.py:28-391 through sdlp.get_bp_updates on the device."""
  for name, fg, evidence, ftype in models.lp_bp_cases():
    sdlp = infer.build_inferer(fg.bp_state, backend="sdlp")
    rng = np.random.RandomState(11)
    arrays = sdlp.init(evidence_updates=evidence,
                       ftov_msgs_updates={ftype: rng.normal(size=fg.bp_state.ftov_msgs.value.shape)})
    models.check_lp_bp_properties(lambda temp: sdlp.get_bp_updates(arrays, temp), sdlp.context, temperature)
