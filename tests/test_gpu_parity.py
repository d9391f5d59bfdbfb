"""GPU parity: the sm_100a kernels behind include/pgx.h against the CPU oracle and
the reference's golden vectors.  Every call goes through the C ABI
(pgmax_b200._native -> libpgx.so).

Tolerances: max-product (T=0) messages are compared at 1e-6 absolute and MAP
decodings must be identical (exactly-tied variables are counted and excluded);
sum-product messages / marginals within 1e-5 absolute (BASELINE.json north_star).
"""

import os

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import infer

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _finite_close(a, b, atol):
  """allclose that treats the -1e32 clip floor as a category, not a number."""
  a, b = np.asarray(a), np.asarray(b)
  floor_a, floor_b = a <= -1e31, b <= -1e31
  np.testing.assert_array_equal(floor_a, floor_b)
  np.testing.assert_allclose(a[~floor_a], b[~floor_b], atol=atol, rtol=1e-5)


def _run_both(bp, arrays, num_iters, damping, temperature):
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(
      graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, num_iters, damping,
      temperature)
  got, got_d = bp.run_with_diffs(arrays, num_iters=num_iters, damping=damping,
                                 temperature=temperature)
  return graph, want, want_d, got, got_d


def test_cut_model_golden():
  gold = np.load(os.path.join(GOLDEN, "e2e_sanity.npz"))
  fg, bp_state, grid_vars, additional_vars = models.cut_model()
  bp = infer.BP(bp_state, temperature=0.0)
  arrays = bp.init()
  out = bp.run(arrays, num_iters=100, damping=0.5)
  np.testing.assert_allclose(out.ftov_msgs, gold["true_final_msgs_output"], atol=1e-6, rtol=1e-5)
  decoded = infer.decode_map_states(bp.get_beliefs(out))
  fused = bp.get_map_states(out)
  groups = {"grid_vars": grid_vars, "additional_vars": additional_vars}
  for name, index, state in zip(gold["map_groups"], gold["map_indices"], gold["map_states"]):
    assert decoded[groups[str(name)]][tuple(index)] == state
    assert fused[groups[str(name)]][tuple(index)] == state
  # deprecated alias gives the same answer (tests/fgraph/test_fgraph.py:249-282)
  with pytest.warns(UserWarning):
    out2 = bp.run_bp(arrays, num_iters=100, damping=0.5)
  np.testing.assert_array_equal(out.ftov_msgs, out2.ftov_msgs)


@pytest.mark.parametrize("temperature", [0.0, 0.5, 1.0])
def test_cut_model_vs_oracle(temperature):
  _, bp_state, _, _ = models.cut_model()
  bp = infer.BP(bp_state, temperature=temperature)
  arrays = bp.init()
  _, want, want_d, got, got_d = _run_both(bp, arrays, 30, 0.5, temperature)
  _finite_close(got.ftov_msgs, want, 1e-6 if temperature == 0.0 else 1e-5)
  np.testing.assert_allclose(got_d, want_d, atol=1e-5)


@pytest.mark.parametrize("temperature", [0.0, 0.05, 1.0])
@pytest.mark.parametrize("batch", [None, 3, 40])
def test_ising_vs_oracle(temperature, batch):
  fg, variables, evidence = models.ising_model(n=12, batch=batch)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={variables: evidence})
  graph, want, want_d, got, got_d = _run_both(bp, arrays, 20, 0.5, temperature)
  atol = 1e-6 if temperature == 0.0 else 1e-5
  np.testing.assert_allclose(got.ftov_msgs, want, atol=atol)
  np.testing.assert_allclose(got_d, want_d, atol=atol)
  if temperature == 0.0:
    # max-product on pairwise binary factors involves no transcendental: bit-exact
    np.testing.assert_array_equal(got.ftov_msgs, want)
  # beliefs / decode / marginals
  want_b = bp_oracle.flat_beliefs(graph, want, arrays.evidence)
  got_b = bp.context.flat_beliefs(got)
  np.testing.assert_allclose(got_b, want_b, atol=atol * 4)
  states, marg, ties = bp.context.decode(got, marginals=True)
  w_states, w_marg, w_ties = bp_oracle.decode_flat(graph, got_b)
  np.testing.assert_array_equal(states, w_states)
  np.testing.assert_array_equal(ties, w_ties)
  np.testing.assert_allclose(marg, w_marg, atol=1e-6)


@pytest.mark.parametrize("num_iters", [20, 200])
def test_rbm24_reference_decodings(num_iters):
  gold = np.load(os.path.join(GOLDEN, "rbm24.npz"))
  for idx in range(0, 50, 7):
    W, bh, bv = gold["W"][idx], gold["bh"][idx], gold["bv"][idx]
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    out = bp.run(bp.init(), num_iters=num_iters, damping=0.5)
    states = bp.get_map_states(out)
    np.testing.assert_array_equal(states[hidden], gold[f"hidden_cpu_{num_iters}"][idx])
    np.testing.assert_array_equal(states[visible], gold[f"visible_cpu_{num_iters}"][idx])


def _small_rbm(nh=12, nv=20, batch=37, temperature=0.0, scale=1.0):
  rng = np.random.default_rng(0)
  W, bh, bv = scale * rng.normal(size=(nh, nv)), rng.logistic(size=nh), rng.logistic(size=nv)
  fg, hidden, visible = models.rbm_model(W, bh, bv)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={
      hidden: rng.gumbel(size=(batch, nh, 2)), visible: rng.gumbel(size=(batch, nv, 2))})
  return bp, arrays


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_rbm_batched_exact_order_vs_oracle(temperature):
  """Two-pass path (serial summation order): max-product is bit-exact with the oracle."""
  bp, arrays = _small_rbm(temperature=temperature)
  bp.context.plan.set_exact_order(True)
  assert arrays.evidence.shape == (37, 2 * (12 + 20))
  _, want, want_d, got, got_d = _run_both(bp, arrays, 25, 0.5, temperature)
  if temperature == 0.0:
    np.testing.assert_array_equal(got.ftov_msgs, want)
  np.testing.assert_allclose(got.ftov_msgs, want, atol=1e-5)
  np.testing.assert_allclose(got_d, want_d, atol=1e-5)


@pytest.mark.parametrize("temperature", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("shape", [(12, 20, 37), (5, 33, 64), (40, 17, 100), (2, 2, 32)])
def test_rbm_fused_single_pass_vs_oracle(temperature, shape):
  """Dense-grid pairwise blocks, batch > 16: one pass per iteration with tiled partial
  sums (tree summation order).  Weak couplings keep BP contractive so that rounding-order
  differences do not amplify; tolerance = north-star 1e-5."""
  nh, nv, batch = shape
  bp, arrays = _small_rbm(nh, nv, batch, temperature, scale=0.15)
  assert bp.context.plan.has_fused_blocks
  _, want, want_d, got, got_d = _run_both(bp, arrays, 25, 0.5, temperature)
  np.testing.assert_allclose(got.ftov_msgs, want, atol=1e-5)
  np.testing.assert_allclose(got_d, want_d, atol=1e-5)
  states, _, _ = bp.context.decode(got)
  graph = bp_oracle.graph_from_context(bp.context)
  w_states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, want, arrays.evidence))
  beliefs = bp_oracle.flat_beliefs(graph, want, arrays.evidence).reshape(batch, -1, 2)
  near_tie = np.abs(beliefs[..., 0] - beliefs[..., 1]) < 1e-4
  assert np.array_equal(states[~near_tie], w_states[~near_tie])
  # strong couplings (BP no longer contractive): only a short horizon is comparable
  bp, arrays = _small_rbm(nh, nv, batch, temperature, scale=1.0)
  _, want, _, got, _ = _run_both(bp, arrays, 3, 0.5, temperature)
  np.testing.assert_allclose(got.ftov_msgs, want, atol=1e-5)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
@pytest.mark.parametrize("shape", [(12, 20, 37), (6, 9, 288), (33, 70, 64)])
def test_rbm_fused_initial_messages_and_resume(temperature, shape):
  """Single-pass path on binary-difference storage, every way messages can enter it:
  (a) resume: run(5) then run(7) on its (batched, normalised) output == run(12) up to the
      summation order of the first variable sums (first iteration reads the full layout, the
      rest the compressed one), and is deterministic;
  (b) non-zero, un-normalised initial messages shared by the batch ([E_s] vector: normalised
      once, broadcast compressed) == the same vector tiled to [B, E_s] (full-layout staging);
  (c) 9 sample tiles (two groups of 8 warps, 7 of them idle) for the batch-288 shape."""
  from pgmax_b200.infer.bp_state import BPArrays
  nh, nv, batch = shape
  bp, arrays = _small_rbm(nh, nv, batch, temperature, scale=0.4)
  assert bp.context.plan.has_fused_blocks
  one, one_d = bp.run_with_diffs(arrays, num_iters=12, damping=0.5, temperature=temperature)
  mid = bp.run(arrays, num_iters=5, damping=0.5, temperature=temperature)
  assert np.asarray(mid.ftov_msgs).shape == (batch, arrays.ftov_msgs.shape[-1])
  two, two_d = bp.run_with_diffs(mid, num_iters=7, damping=0.5, temperature=temperature)
  # (a resumed run forms its first variable sums serially from the messages, the continuous run
  # from the previous iteration's tiled partial sums: same values up to summation order)
  np.testing.assert_allclose(one.ftov_msgs, two.ftov_msgs, atol=1e-4)
  np.testing.assert_allclose(np.asarray(one_d)[:, 5:], two_d, atol=1e-4)
  again, again_d = bp.run_with_diffs(mid, num_iters=7, damping=0.5, temperature=temperature)
  np.testing.assert_array_equal(two.ftov_msgs, again.ftov_msgs)  # deterministic
  np.testing.assert_array_equal(two_d, again_d)
  rng = np.random.default_rng(3)
  init = rng.normal(size=arrays.ftov_msgs.shape[-1]).astype(np.float32)
  shared = BPArrays(log_potentials=arrays.log_potentials, ftov_msgs=init, evidence=arrays.evidence)
  tiled = BPArrays(log_potentials=arrays.log_potentials, ftov_msgs=np.tile(init, (batch, 1)),
                   evidence=arrays.evidence)
  a = bp.run(shared, num_iters=6, damping=0.5, temperature=temperature)
  b = bp.run(tiled, num_iters=6, damping=0.5, temperature=temperature)
  np.testing.assert_array_equal(a.ftov_msgs, b.ftov_msgs)
  # and against the serial-order two-pass path (tree vs serial sums: tolerance, short horizon)
  bp.context.plan.set_exact_order(True)
  c = bp.run(shared, num_iters=6, damping=0.5, temperature=temperature)
  bp.context.plan.set_exact_order(False)
  np.testing.assert_allclose(a.ftov_msgs, c.ftov_msgs, atol=1e-4)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_rbm_fused_half_batch_pipeline_is_bit_identical(temperature):
  """>= 16 sample tiles: the two halves of the batch run as two pipelined chains on two streams
  (PGX_PATH_HALF_BATCH).  Samples are independent, so every message and delta must equal the
  single-chain run bit for bit; 17 tiles = halves of 8 and 9 tiles, the last one partial."""
  bp, arrays = _small_rbm(6, 9, 530, temperature, scale=0.4)
  plan = bp.context.plan
  got, got_d = bp.run_with_diffs(arrays, num_iters=9, damping=0.5, temperature=temperature)
  plan.disable_paths(plan.PATH_HALF_BATCH)
  ref, ref_d = bp.run_with_diffs(arrays, num_iters=9, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  one = bp.run(arrays, num_iters=1, damping=0.5, temperature=temperature)
  plan.disable_paths(plan.PATH_HALF_BATCH)
  one_ref = bp.run(arrays, num_iters=1, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  np.testing.assert_array_equal(one.ftov_msgs, one_ref.ftov_msgs)


def test_rbm_full_size_fused_properties():
  """BASELINE configs[1] shape (RBM 784 x 500) at batch 64 / 96, sum-product:
  (1) the single-pass path agrees with the serial-order two-pass path over a short horizon
      (var sums here are sums of ~800 terms of magnitude ~1e2: one fp32 ulp is 3e-5, and BP on
      this model is chaotic, so only a short horizon is comparable at all);
  (2) sharding invariance: a sample's result does not depend on its batch-mates (bit-exact);
  (3) messages stay normalised: every edge has max 0; deltas are finite."""
  rs = np.random.RandomState(0)
  nh, nv = 500, 784
  W, bh, bv = rs.normal(size=(nh, nv)), rs.logistic(size=nh), rs.logistic(size=nv)
  fg, hidden, visible = models.rbm_model(W, bh, bv)
  bp = infer.BP(fg.bp_state, temperature=1.0)
  rng = np.random.default_rng(0)
  ev_h, ev_v = rng.gumbel(size=(96, nh, 2)), rng.gumbel(size=(96, nv, 2))
  big = bp.init(evidence_updates={hidden: ev_h, visible: ev_v})
  small = bp.init(evidence_updates={hidden: ev_h[:64], visible: ev_v[:64]})
  plan = bp.context.plan
  assert plan.has_fused_blocks
  fused2 = bp.run(small, num_iters=2, damping=0.5).ftov_msgs
  plan.set_exact_order(True)
  exact2 = bp.run(small, num_iters=2, damping=0.5).ftov_msgs
  plan.set_exact_order(False)
  assert np.max(np.abs(fused2 - exact2)) < 2e-4
  out_small, d_small = bp.run_with_diffs(small, num_iters=6, damping=0.5)
  out_big, d_big = bp.run_with_diffs(big, num_iters=6, damping=0.5)
  np.testing.assert_array_equal(out_big.ftov_msgs[:64], out_small.ftov_msgs)
  np.testing.assert_array_equal(d_big[:64], d_small)
  edge_max = out_big.ftov_msgs.reshape(96, -1, 2).max(axis=-1)
  assert np.all(edge_max == 0.0) and np.all(np.isfinite(d_big))


TEMPS = [(0.0, 1e-5), (0.001, 5e-3), (0.3, 5e-3), (0.8, 1e-5)]


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
@pytest.mark.parametrize("seed", range(8))
def test_logical_vs_oracle_and_vs_enum(kind, seed):
  """The reference's differential test (tests/factor/test_or.py:30-290 etc.) on the GPU:
  logical factors == equivalent EnumFactors, and both == the oracle."""
  temperature, atol = TEMPS[seed % 4]
  data = models.logical_pair(kind, seed)
  beliefs = []
  for entry in data["graphs"]:
    bp = infer.BP(entry[0].bp_state, temperature=temperature)
    arrays = models.init_logical(bp, entry, data)
    graph, want, _, got, _ = _run_both(bp, arrays, 5, 0.5, temperature)
    # vs the oracle: same branch structure, so tight even at low temperature
    _finite_close(got.ftov_msgs, want, 2e-5 if temperature >= 0.5 or temperature == 0 else 1e-3)
    beliefs.append(bp.context.flat_beliefs(got))
  np.testing.assert_allclose(beliefs[0], beliefs[1], atol=atol, rtol=0)


def test_clipping_equality_factors():
  """tests/test_clipping.py:24-58: messages hit the -1e32 floor; all decode to 1."""
  from pgmax_b200 import fgraph, fgroup, vgroup
  num = 10
  variables = vgroup.NDVarArray(num_states=2, shape=(num,))
  fg = fgraph.FactorGraph(variables)
  pairs = [[variables[i], variables[j]] for i in range(num) for j in range(i + 1, num)]
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=pairs, factor_configs=np.array([[0, 0], [1, 1]])))
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates={variables: np.tile(np.array([0.0, 1.0]), (num, 1))})
  graph, want, _, got, _ = _run_both(bp, arrays, 100, 0.0, 0.0)
  _finite_close(got.ftov_msgs, want, 1e-6)
  np.testing.assert_array_equal(bp.get_map_states(got)[variables], np.ones(num))


def test_infinite_potentials_are_clipped_inside_run():
  """tests/test_energy.py:75-100: -inf potentials are clipped to -1e6 in run only."""
  from pgmax_b200 import factor, fgraph, vgroup
  variables = vgroup.NDVarArray(num_states=2, shape=(2,))
  fg = fgraph.FactorGraph(variables)
  fg.add_factors(factor.EnumFactor(
      variables=[variables[0], variables[1]],
      factor_configs=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]),
      log_potentials=np.array([-np.inf, -np.inf, -np.inf, 0.0])))
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init()
  out = bp.run(arrays, num_iters=1, damping=0.0)
  assert out.log_potentials is arrays.log_potentials  # returned unclipped
  np.testing.assert_array_equal(bp.get_map_states(out)[variables], [1, 1])
  graph = bp_oracle.graph_from_context(bp.context)
  want, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 1, 0.0, 0.0)
  np.testing.assert_array_equal(out.ftov_msgs, want)


def test_ragged_num_states():
  """tests/fgraph/test_fgraph.py:285-333: variables with different numbers of states."""
  from pgmax_b200 import factor, fgraph, vgroup
  rng = np.random.default_rng(1)
  ns = np.array([2, 3, 4, 3])
  variables = vgroup.NDVarArray(num_states=ns, shape=(4,))
  fg = fgraph.FactorGraph(variables)
  for a, b in [(0, 1), (1, 2), (2, 3), (3, 0)]:
    cfg = np.array([[i, j] for i in range(ns[a]) for j in range(ns[b]) if (i + j) % 3 != 2])
    fg.add_factors(factor.EnumFactor(variables=[variables[a], variables[b]], factor_configs=cfg,
                                     log_potentials=rng.normal(size=cfg.shape[0])))
  for temperature in (0.0, 1.0):
    bp = infer.BP(fg.bp_state, temperature=temperature)
    arrays = bp.init(evidence_updates={variables[i]: rng.gumbel(size=ns[i]) for i in range(4)})
    graph, want, _, got, _ = _run_both(bp, arrays, 10, 0.5, temperature)
    _finite_close(got.ftov_msgs, want, 1e-5)
    beliefs = bp.get_beliefs(got)[variables]
    assert beliefs.shape == (4, 4) and np.isneginf(beliefs[0, 2:]).all()
    marg = infer.get_marginals(bp.get_beliefs(got))[variables]
    np.testing.assert_allclose(marg.sum(-1), 1.0, atol=1e-6)


def test_infer_host_end_to_end():
  fg, variables, evidence = models.ising_model(n=10, batch=5)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates={variables: evidence})
  res = bp.infer_host(arrays, num_iters=30, damping=0.5, marginals=True, return_msgs=True)
  got = bp.run(arrays, num_iters=30, damping=0.5)
  np.testing.assert_array_equal(res["ftov_msgs"], got.ftov_msgs)
  states, marg, ties = bp.context.decode(got, marginals=True)
  np.testing.assert_array_equal(res["flat_map_states"], states)
  np.testing.assert_array_equal(res["tie_counts"], ties)
  np.testing.assert_allclose(res["marginals"], marg, atol=0)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_infer_host_decodes_from_the_final_variable_sums(temperature):
  """pgx_infer_host without a message output (the benchmark's end-to-end call): the run leaves the
  variable sums of the final messages in the workspace (PGX_RUN_FINAL_SUMS), the decode reads
  those and the messages are never written in the ABI layout (PGX_RUN_SKIP_OUTPUT).  Generic
  batched path: identical to decoding the run's messages (same serial sums); single-pass RBM path
  (tree-order partial sums): identical MAP away from near-ties, marginals to 1e-5; one sample and
  OR / AND graphs fall back to decoding from the messages."""
  fg, variables, evidence = models.ising_model(n=10, batch=40)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={variables: evidence})
  for _ in range(3):  # direct, captured, replayed
    res = bp.infer_host(arrays, num_iters=12, damping=0.5, marginals=True)
    got = bp.run(arrays, num_iters=12, damping=0.5)
    states, marg, ties = bp.context.decode(got, marginals=True)
    np.testing.assert_array_equal(res["flat_map_states"], states)
    np.testing.assert_array_equal(res["tie_counts"], ties)
    np.testing.assert_array_equal(res["marginals"], marg)
    assert res["ftov_msgs"] is None
  bp, arrays = _small_rbm(9, 14, 70, temperature, scale=0.3)
  assert bp.context.plan.has_fused_blocks
  res = bp.infer_host(arrays, num_iters=10, damping=0.5, marginals=True)
  got = bp.run(arrays, num_iters=10, damping=0.5, temperature=temperature)
  states, marg, _ = bp.context.decode(got, marginals=True)
  np.testing.assert_allclose(res["marginals"], marg, atol=1e-5)
  beliefs = np.asarray(bp.context.flat_beliefs(got)).reshape(70, -1, 2)
  clear = np.abs(beliefs[..., 0] - beliefs[..., 1]) > 1e-4
  assert np.array_equal(res["flat_map_states"][clear], states[clear])
  one, _, ev1 = models.ising_model(n=6)
  bp1 = infer.BP(one.bp_state, temperature=temperature)
  a1 = bp1.init(evidence_updates={_: ev1})
  r1 = bp1.infer_host(a1, num_iters=8, damping=0.5)
  s1, _, _ = bp1.context.decode(bp1.run(a1, num_iters=8, damping=0.5))
  np.testing.assert_array_equal(r1["flat_map_states"], s1)


def test_split_run_equals_single_run():
  """Resume contract: run(a)+run(b) == run(a+b) because NC is idempotent."""
  fg, variables, evidence = models.ising_model(n=8)
  bp = infer.BP(fg.bp_state, temperature=1.0)
  arrays = bp.init(evidence_updates={variables: evidence})
  one = bp.run(arrays, num_iters=12, damping=0.5)
  two = bp.run(bp.run(arrays, num_iters=5, damping=0.5), num_iters=7, damping=0.5)
  np.testing.assert_array_equal(one.ftov_msgs, two.ftov_msgs)


@pytest.mark.parametrize("temperature", [0.0, 0.05])
@pytest.mark.parametrize("batch", [None, 40])
def test_pull_mode_streaming_equals_persistent(temperature, batch):
  """Pairwise-only graphs on low-degree variables run without a var-sum array (pull mode):
  small ones as ONE cooperative launch for all iterations, otherwise one launch per
  iteration.  Enabling the profiling hooks forces the per-iteration launches; both must be
  bit-identical (and both are compared with the oracle in test_ising_vs_oracle)."""
  fg, variables, evidence = models.ising_model(n=10, batch=batch)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={variables: evidence})
  persistent, d_p = bp.run_with_diffs(arrays, num_iters=40, damping=0.5)
  plan = bp.context.plan
  plan.disable_paths(plan.PATH_LATTICE)  # a one-sample grid would otherwise stream through k_lattice
  plan.profile_enable(True)
  streaming, d_s = bp.run_with_diffs(arrays, num_iters=40, damping=0.5)
  launches, _, name = plan.profile_read()
  plan.profile_enable(False)
  assert launches == 40 and name == "k_enum_pw2_pull"
  np.testing.assert_array_equal(persistent.ftov_msgs, streaming.ftov_msgs)
  np.testing.assert_array_equal(d_p, d_s)
  if batch is None:  # and the index-free lattice kernel, bit for bit
    plan.disable_paths(plan.PATH_RESIDENT | plan.PATH_PULL)
    assert plan.is_lattice
    lattice, d_l = bp.run_with_diffs(arrays, num_iters=40, damping=0.5)
    np.testing.assert_array_equal(persistent.ftov_msgs, lattice.ftov_msgs)
    np.testing.assert_array_equal(d_p, d_l)
    plan.disable_paths(0)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_rcn_shaped_merged_single_factor_groups(temperature):
  """RCN-shaped graph (examples/rcn.ipynb cell 26): every factor is its own EnumFactorGroup
  but only a few distinct config tables exist; the plan merges them into one launch per table.
  25-state variables -> 50 edge-states per factor (k_enum_small) and 81-state variables ->
  162 edge-states (k_enum_big, shared-memory CTA kernel)."""
  for hps, vps in ((2, 2), (4, 4)):
    fg, groups, evidence = models.rcn_model(num_models=3, num_vars=7, hps=hps, vps=vps, radii=(1, 2),
                                            extra_edges=4, seed=1)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    arrays = bp.init(evidence_updates=evidence)
    graph, want, want_d, got, got_d = _run_both(bp, arrays, 10, 0.5, temperature)
    _finite_close(got.ftov_msgs, want, 1e-6 if temperature == 0.0 else 1e-5)
    np.testing.assert_allclose(got_d, want_d, atol=1e-5)
    states, _, _ = bp.context.decode(got)
    w_states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, got.ftov_msgs, arrays.evidence))
    np.testing.assert_array_equal(states, w_states)


def test_deconvolution_logical_batched():
  """Binary deconvolution (AND + OR factor groups, examples/pmp_binary_deconvolution.ipynb),
  small image, batch of 5, max-product: GPU vs oracle."""
  fg, groups = models.deconv_model(im_height=8, im_width=8, n_feat=2, feat_height=3, feat_width=3)
  evidence = models.deconv_evidence(groups, batch=5)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  _, want, want_d, got, got_d = _run_both(bp, arrays, 15, 0.5, 0.0)
  _finite_close(got.ftov_msgs, want, 1e-4)  # messages reach |logit(1e-100)| = 230: ulp 1.5e-5
  np.testing.assert_allclose(got_d, want_d, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("kind", ["or", "and"])
@pytest.mark.parametrize("parents_range", [(1, 2), (3, 4), (1, 5)])
@pytest.mark.parametrize("temperature", [0.0, 0.8])
def test_logical_narrow_and_uniform_groups(kind, parents_range, temperature):
  """Groups whose factors all have the same small number of parents (k_logical_uniform: 1 and
  3 parents here, 2 in the deconvolution test) and ragged groups of <= 4 parents
  (k_logical<narrow>), against the oracle and against the equivalent EnumFactors."""
  data = models.logical_pair(kind, 2, parents_range=parents_range)
  beliefs = []
  for entry in data["graphs"]:
    bp = infer.BP(entry[0].bp_state, temperature=temperature)
    arrays = models.init_logical(bp, entry, data)
    _, want, _, got, _ = _run_both(bp, arrays, 5, 0.5, temperature)
    _finite_close(got.ftov_msgs, want, 2e-5)
    beliefs.append(bp.context.flat_beliefs(got))
  np.testing.assert_allclose(beliefs[0], beliefs[1], atol=1e-5, rtol=0)
