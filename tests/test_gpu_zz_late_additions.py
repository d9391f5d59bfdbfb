"""GPU tests added after the round's last GPU session (validated against the oracle on CPU only; the file
name makes them run last).

One EnumFactorGroup over variables with different numbers of states (split into uniform
blocks by the host mirror, pgmax_b200/factor/enum.py compile_wiring) against the oracle.  Runs
last in the suite (file name): added after the round's last GPU session, validated on CPU only
(the device sees nothing new - the blocks look like separate factor groups)."""

import numpy as np
import pytest

from oracle import bp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("temperature", [0.0, 1.0])
@pytest.mark.parametrize("batch", [None, 3])
def test_ragged_enum_factor_group(temperature, batch):
  num_states = np.array([2, 2, 3, 3, 3, 4, 2])
  variables = vgroup.NDVarArray(num_states=num_states, shape=(7,))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = [(0, 1), (0, 6), (2, 3), (3, 4), (5, 2), (1, 6)]
  rng = np.random.RandomState(0)
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[variables[a], variables[b]] for a, b in pairs],
      factor_configs=np.array([[0, 0], [1, 1], [0, 1]]), log_potentials=rng.normal(size=(len(pairs), 3))))
  bp = infer.BP(fg.bp_state, temperature=temperature)
  shape = (7, 4) if batch is None else (batch, 7, 4)
  arrays = bp.init(evidence_updates={variables: rng.gumbel(size=shape)})
  got, got_d = bp.run_with_diffs(arrays, num_iters=8, damping=0.5, temperature=temperature)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 8, 0.5,
                                          temperature)
  atol = 1e-6 if temperature == 0.0 else 1e-5
  got_m, want_m = np.asarray(got.ftov_msgs), np.asarray(want)
  floor = want_m <= -1e31
  np.testing.assert_array_equal(got_m <= -1e31, floor)
  np.testing.assert_allclose(got_m[~floor], want_m[~floor], atol=atol)
  np.testing.assert_allclose(got_d, want_d, atol=atol)


@pytest.mark.parametrize("n_units", [40, 100, 200])
def test_rbm_large_reference_decodings(n_units):
  """The reference's stored decodings of 40-, 100- and 200-unit RBMs after 20 max-product
  iterations (tests/golden/rbm_large.npz, benchmark/precomputed_results/): one sample, serial
  summation order - the device decodes the same states (as it does for the 24-unit RBMs in
  test_gpu_parity.py::test_rbm24_reference_decodings)."""
  import os
  import models

  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbm_large.npz"))
  for idx in range(4):
    W, bh, bv = (gold[f"{k}_{n_units}_{idx}"] for k in ("W", "bh", "bv"))
    fg, hidden, visible = models.rbm_model(W, bh, bv)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    out = bp.run(bp.init(), num_iters=20, damping=0.5)
    states = bp.get_map_states(out)
    np.testing.assert_array_equal(states[hidden], gold[f"hidden_cpu_{n_units}_{idx}"])
    np.testing.assert_array_equal(states[visible], gold[f"visible_cpu_{n_units}_{idx}"])


def _partition_worker(rank, world, port, temperature, iters, out, which="cut"):
  import os
  import torch
  import torch.distributed as dist
  import models
  from pgmax_b200 import dist as pdist

  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    if which == "cut":
      fg, bp_state, _, _ = models.cut_model()
      bp = infer.BP(bp_state, temperature=temperature)
      arrays = bp.init()
    else:  # OR + AND network on shared leaves (tests/test_gpu_logical_pull.py)
      import test_gpu_logical_pull
      fg, groups = test_gpu_logical_pull.random_logical_network(1)
      bp_state = fg.bp_state
      bp = infer.BP(bp_state, temperature=temperature)
      rng = np.random.default_rng(2)
      arrays = bp.init(evidence_updates={g: rng.gumbel(size=g.shape + (2,)) * 2.0 for g in groups.values()})
    flat = pdist.flat_from_state(bp_state.fg_state)
    part = pdist.partition_flat(flat, world, rank)
    dev = f"cuda:{rank}"
    runner = pdist.PartitionRunner(part, pdist.PgxStepEngine(part.flat, dev), dev)
    lp = np.asarray(arrays.log_potentials, np.float32)
    ev = np.asarray(arrays.evidence, np.float32)
    msgs = runner.run(lp[part.potential_index], ev[part.var_state_index], iters, 0.5, temperature)
    graph = bp_oracle.graph_from_context(bp.context)
    want, _ = bp_oracle.run_bp(graph, lp, arrays.ftov_msgs, ev, iters, 0.5, temperature)
    out[rank] = float(np.max(np.abs(msgs.cpu().numpy() - want[part.msg_index])))
  finally:
    dist.destroy_process_group()


def _gpu_count():
  import torch
  return torch.cuda.device_count()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_factor_partition_two_gpus_match_single_graph(temperature):
  """pdist.partition_flat / PartitionRunner over NCCL on 2 GPUs (the gloo / oracle-engine version
  of this test runs on CPU in tests/test_dist_gloo.py)."""
  import socket
  import torch.multiprocessing as mp

  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = mp.Manager().dict()
  mp.spawn(_partition_worker, args=(2, port, temperature, 10, out), nprocs=2, join=True)
  assert max(out.values()) <= 1e-5, dict(out)
  if temperature == 0.0:  # OR / AND factors cut across the two ranks
    out = mp.Manager().dict()
    mp.spawn(_partition_worker, args=(2, port, temperature, 10, out, "logical"), nprocs=2, join=True)
    assert max(out.values()) <= 2e-5, dict(out)


@pytest.mark.parametrize("temperature", [0.0])
def test_half_batch_pipeline_with_merged_max_product_blocks(temperature):
  """A dense pairwise-binary grid (single-pass kernel, half-batch pipeline at >= 16 sample tiles)
  in the same graph as large sorted pairwise EnumFactors (merged max-product launch, work units
  handed out through an atomic counter): the two half-batch chains run that launch concurrently,
  each on its own counter.  Pipeline on == pipeline off, bit for bit, and == the oracle for a few
  samples of both halves."""
  rng = np.random.RandomState(0)
  rows = vgroup.NDVarArray(num_states=2, shape=(6,))
  cols = vgroup.NDVarArray(num_states=2, shape=(9,))
  wide = vgroup.NDVarArray(num_states=40, shape=(4,))
  fg = fgraph.FactorGraph(variable_groups=[rows, cols, wide])
  lpm = np.zeros((54, 2, 2))
  lpm[:, 1, 1] = 0.4 * rng.normal(size=54)
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[rows[i], cols[j]] for i in range(6) for j in range(9)], log_potential_matrix=lpm))
  configs = np.stack(np.meshgrid(np.arange(40), np.arange(40), indexing="ij"), axis=-1).reshape(-1, 2)
  configs = configs[np.abs(configs[:, 0] - configs[:, 1]) <= 9]  # banded, sorted by the first state
  pairs = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[wide[a], wide[b]] for a, b in pairs], factor_configs=configs,
      log_potentials=rng.normal(size=(len(pairs), configs.shape[0]))))
  bp = infer.BP(fg.bp_state, temperature=temperature)
  batch = 530
  arrays = bp.init(evidence_updates={rows: rng.gumbel(size=(batch, 6, 2)), cols: rng.gumbel(size=(batch, 9, 2)),
                                     wide: rng.gumbel(size=(batch, 4, 40))})
  plan = bp.context.plan
  assert plan.has_fused_blocks
  got, got_d = bp.run_with_diffs(arrays, num_iters=7, damping=0.5, temperature=temperature)
  plan.disable_paths(plan.PATH_HALF_BATCH)
  ref, ref_d = bp.run_with_diffs(arrays, num_iters=7, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  graph = bp_oracle.graph_from_context(bp.context)
  for b in (0, 255, 256, 529):
    want, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence[b], 7, 0.5, temperature)
    np.testing.assert_allclose(np.asarray(got.ftov_msgs)[b], want, atol=1e-5)


def test_batched_beliefs_and_map_states_with_a_vardict():
  """The batch axis through get_beliefs / get_map_states on a graph that holds a VarDict
  (VarDict.unflatten keeps the leading batch axis, as NDVarArray.unflatten does)."""
  rng = np.random.RandomState(1)
  names = ("a", "b", "c")
  vd = vgroup.VarDict(variable_names=names, num_states=3)
  arr = vgroup.NDVarArray(num_states=3, shape=(2,))
  fg = fgraph.FactorGraph(variable_groups=[vd, arr])
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[vd["a"], vd["b"]], [vd["b"], vd["c"]], [vd["c"], arr[0]], [arr[0], arr[1]]],
      log_potential_matrix=rng.normal(size=(4, 3, 3))))
  bp = infer.BP(fg.bp_state, temperature=0.0)
  ev_arr = rng.gumbel(size=(5, 2, 3))
  arrays = bp.init(evidence_updates={arr: ev_arr})
  out = bp.run(arrays, num_iters=10, damping=0.5)
  beliefs = bp.get_beliefs(out)
  states = bp.get_map_states(out)
  assert beliefs[arr].shape == (5, 2, 3) and states[arr].shape == (5, 2)
  for name in names:
    assert beliefs[vd][name].shape == (5, 3) and states[vd][name].shape == (5,)
  for b in range(5):
    one = bp.init(evidence_updates={arr: ev_arr[b]})
    one_out = bp.run(one, num_iters=10, damping=0.5)
    one_beliefs, one_states = bp.get_beliefs(one_out), bp.get_map_states(one_out)
    for name in names:
      np.testing.assert_array_equal(beliefs[vd][name][b], one_beliefs[vd][name])
      assert states[vd][name][b] == one_states[vd][name]


def test_round_ordered_potentials_are_reused_only_while_unchanged():
  """RCN-size factors: the merged max-product launch reads a round-ordered copy of the potentials
  made once per run; a repeated run on the SAME potentials (same immutable host array, or the
  same device tensor with an unchanged version counter) reuses it
  (PGX_RUN_POTENTIALS_UNCHANGED), modified potentials do not.  Every result == the oracle."""
  import torch
  import models
  from pgmax_b200.infer.bp_state import BPArrays

  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=5, radii=(2, 5), extra_edges=2, seed=7)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  rng = np.random.default_rng(0)
  lp1 = rng.normal(size=arrays.log_potentials.shape).astype(np.float32)
  lp2 = rng.normal(size=arrays.log_potentials.shape).astype(np.float32)
  graph = bp_oracle.graph_from_context(bp.context)
  plan = bp.context.plan
  want = {}
  for name, lp in (("lp1", lp1), ("lp2", lp2)):
    want[name], _ = bp_oracle.run_bp(graph, lp, arrays.ftov_msgs, arrays.evidence, 4, 0.5, 0.0)
  # host arrays: same object twice (second run: cached device copy, no permute launch), then new values
  a1 = BPArrays(log_potentials=lp1, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  n0 = plan.launch_count
  np.testing.assert_array_equal(bp.run(a1, num_iters=4, damping=0.5).ftov_msgs, want["lp1"])
  first = plan.launch_count - n0
  n0 = plan.launch_count
  np.testing.assert_array_equal(bp.run(a1, num_iters=4, damping=0.5).ftov_msgs, want["lp1"])
  assert plan.launch_count - n0 == first - 1            # the permute pass is skipped
  a2 = BPArrays(log_potentials=lp2, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  np.testing.assert_array_equal(bp.run(a2, num_iters=4, damping=0.5).ftov_msgs, want["lp2"])
  # device tensor modified in place between runs: the version counter changes, the copy is redone
  t = torch.from_numpy(lp1.copy()).cuda()
  d1 = BPArrays(log_potentials=t, ftov_msgs=torch.from_numpy(np.asarray(arrays.ftov_msgs).copy()).cuda(),
                evidence=torch.from_numpy(np.asarray(arrays.evidence).copy()).cuda())
  for _ in range(2):
    np.testing.assert_array_equal(bp.run(d1, num_iters=4, damping=0.5).ftov_msgs.cpu().numpy(), want["lp1"])
  t.copy_(torch.from_numpy(lp2))
  for _ in range(3):
    np.testing.assert_array_equal(bp.run(d1, num_iters=4, damping=0.5).ftov_msgs.cpu().numpy(), want["lp2"])


@pytest.mark.parametrize("temperature", [0.0, 0.7])
@pytest.mark.parametrize("batch", [None, 6, 70])
def test_unary_enum_closed_form_is_bit_identical(temperature, batch):
  """k_enum_unary (one-variable EnumFactors over all states: f = ((0 + q) + lp) - q) against the
  general small-factor kernel (PATH_ENUM_UNARY disabled): the same operations, bit-identical, for
  2-state and 5-state unary factors (potentials beyond the +-1e6 clip included) beside pairwise
  ones, one sample / small batch / single-pass RBM path; and against the oracle."""
  import models
  rng = np.random.RandomState(5)
  a = vgroup.NDVarArray(num_states=2, shape=(6,))
  b = vgroup.NDVarArray(num_states=5, shape=(4,))
  fg = fgraph.FactorGraph(variable_groups=[a, b])
  lp_a = rng.normal(size=(6, 2))
  lp_a[2, 1] = 3e6
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=[[a[i]] for i in range(6)],
                                        factor_configs=np.arange(2)[:, None], log_potentials=lp_a))
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=[[b[i]] for i in range(4)],
                                        factor_configs=np.arange(5)[:, None], log_potentials=rng.normal(size=(4, 5))))
  fg.add_factors(fgroup.PairwiseFactorGroup(variables_for_factors=[[a[i], b[i % 4]] for i in range(6)],
                                            log_potential_matrix=rng.normal(size=(6, 2, 5))))
  graphs = [(fg, {a: (6, 2), b: (4, 5)})]
  W, bh, bv = 0.4 * rng.normal(size=(5, 7)), rng.logistic(size=5), rng.logistic(size=7)
  rbm, hidden, visible = models.rbm_model(W, bh, bv)
  graphs.append((rbm, {hidden: (5, 2), visible: (7, 2)}))
  lead = () if batch is None else (batch,)
  for graph_fg, shapes in graphs:
    bp = infer.BP(graph_fg.bp_state, temperature=temperature)
    arrays = bp.init(evidence_updates={vg: rng.gumbel(size=lead + shp) for vg, shp in shapes.items()})
    plan = bp.context.plan
    got, got_d = bp.run_with_diffs(arrays, num_iters=8, damping=0.5, temperature=temperature)
    plan.disable_paths(plan.PATH_ENUM_UNARY)
    ref, ref_d = bp.run_with_diffs(arrays, num_iters=8, damping=0.5, temperature=temperature)
    plan.disable_paths(0)
    np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
    np.testing.assert_array_equal(got_d, ref_d)
    graph = bp_oracle.graph_from_context(bp.context)
    want, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 8, 0.5, temperature)
    np.testing.assert_allclose(got.ftov_msgs, want, atol=1e-6 if temperature == 0.0 else 1e-5)


@pytest.mark.parametrize("temperature", [1.0, 0.3])
@pytest.mark.parametrize("batch", [None, 3])
@pytest.mark.parametrize("extreme", [False, True])
def test_rcn_size_factors_sum_product_one_pass(temperature, batch, extreme):
  """k_enum_big_sumprod_all: sum-product (T > 0) of RCN-size sorted pairwise factors on the round
  schedule of the max-product launch - one visit per configuration, running (max, sum) pairs
  ("online" logsumexp) instead of the reference's max-then-sum (pgmax/factor/enum.py:451-475).
  Random potentials; `extreme`: some beyond the +-1e6 clip and some -inf (states left without a
  finite configuration included) - messages of magnitude 1e6 there, so that case is held to the
  fp32 oracle relatively; the plain case absolutely, judged by the fp64 oracle.  Also against
  the generic thread-per-edge-state kernel (PATH_MERGED_MAX disabled); replayed from the CUDA
  graph bit for bit."""
  import models
  from pgmax_b200.infer.bp_state import BPArrays

  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=6, radii=(2, 4, 7), extra_edges=2, seed=11)
  rng = np.random.default_rng(3)
  if batch is not None:
    evidence = {vg: np.stack([np.where(rng.random(ev.shape) < 0.05, 1.0, -1.0) for _ in range(batch)])
                for vg, ev in evidence.items()}
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=evidence)
  lp = (rng.normal(size=arrays.log_potentials.shape) * 2.0).astype(np.float32)
  if extreme:
    lp[rng.integers(0, lp.size, size=50)] = 3e6
    lp[rng.integers(0, lp.size, size=50)] = -np.inf
  arrays = BPArrays(log_potentials=lp, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  plan = bp.context.plan
  iters = 6
  before = plan.launch_count
  got, got_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  merged = plan.launch_count - before
  plan.disable_paths(plan.PATH_MERGED_MAX)
  before = plan.launch_count
  ref, ref_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  per_group = plan.launch_count - before
  plan.disable_paths(0)
  assert merged < per_group                  # one launch for all 7 factor groups (+ the permute pass)
  graph = bp_oracle.graph_from_context(bp.context)
  want, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5,
                                     temperature)
  with bp_oracle.precision(np.float64):
    exact, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5,
                                        temperature)
  got_m, ref_m = np.asarray(got.ftov_msgs, np.float64), np.asarray(ref.ftov_msgs, np.float64)
  want_m, exact_m = np.asarray(want, np.float64).reshape(got_m.shape), np.asarray(exact, np.float64).reshape(got_m.shape)
  floor = exact_m <= -1e31
  np.testing.assert_array_equal(got_m <= -1e31, floor)
  err_got = np.abs(got_m - exact_m)[~floor].max()
  err_ref = np.abs(ref_m - exact_m)[~floor].max()
  err_oracle = np.abs(want_m - exact_m)[~floor].max()
  print(f"vs fp64: one-pass kernel {err_got:.3g}, generic kernel {err_ref:.3g}, fp32 oracle {err_oracle:.3g}")
  if extreme:
    np.testing.assert_allclose(got_m[~floor], want_m[~floor], rtol=2e-6, atol=4e-5)
    assert err_got <= 2.0 * err_oracle, (err_got, err_oracle)
  else:
    assert err_got <= max(2e-5, 2.0 * err_oracle), (err_got, err_oracle)
    np.testing.assert_allclose(got_m, want_m, atol=4e-5)
  for _ in range(2):                         # capture, replay
    again, again_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
    np.testing.assert_array_equal(again.ftov_msgs, got.ftov_msgs)
    np.testing.assert_array_equal(again_d, got_d)


@pytest.mark.parametrize("temperature", [0.0, 1.0, 0.4])
@pytest.mark.parametrize("batch,lp_batched", [(None, False), (5, False), (70, False), (5, True), (70, True)])
def test_small_enum_configuration_major_walk(temperature, batch, lp_batched):
  """k_enum_small_cm (small EnumFactors, <= 32 edge-states: one walk over the configurations per
  pass, per-thread shared-memory columns) against k_enum_small (PATH_ENUM_CONFIG_MAJOR disabled:
  per-edge-state list walks) and the oracle, on 17 x 3-state pairwise factors (the reference's
  "heretic" test model, tests/test_pgmax.py:424-475, cut to 8 x 8 hidden variables), ragged
  three-variable factors with a sparse configuration table (edge-states in one configuration only
  / in none), a four-variable factor (run-time arity instantiation), potentials beyond the clip,
  shared by the batch (staged per warp by k_enum_pair_dense) or per sample (`lp_batched`).
  Max-product: bit-identical.  Sum-product: ex2 instead of expf, same order of additions - at the
  sum-product tolerance, judged by fp64."""
  from pgmax_b200.infer.bp_state import BPArrays
  rng = np.random.default_rng(2)
  pixels = vgroup.NDVarArray(shape=(10, 10), num_states=3)
  hidden = vgroup.NDVarArray(shape=(8, 8), num_states=17)
  extra = vgroup.NDVarArray(shape=(6,), num_states=np.array([2, 4, 3, 2, 4, 3]))
  fg = fgraph.FactorGraph([pixels, hidden, extra])
  for dr in range(3):
    for dc in range(3):
      fg.add_factors(fgroup.PairwiseFactorGroup(
          variables_for_factors=[[hidden[r, c], pixels[r + dr, c + dc]] for r in range(8) for c in range(8)],
          log_potential_matrix=rng.normal(size=(17, 3))))
  # the same complete tables with the few-state variable FIRST, and 2- / 4-state few sides (k_enum_pair_few)
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[pixels[r, c], hidden[(r + 4) % 8, (c + 4) % 8]] for r in range(8) for c in range(8)],
      log_potential_matrix=rng.normal(size=(3, 17))))
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[hidden[r, 0], extra[0]] for r in range(8)], log_potential_matrix=rng.normal(size=(17, 2))))
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[extra[1], hidden[0, c]] for c in range(8)], log_potential_matrix=rng.normal(size=(4, 17))))
  # sparse three-variable table over (2, 4, 3) states: state 3 of the middle variable is in no configuration
  configs = np.array([[0, 0, 0], [0, 1, 2], [1, 2, 1], [1, 0, 2], [0, 2, 0], [1, 1, 1]])
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[extra[0], extra[1], extra[2]], [extra[3], extra[4], extra[5]]],
      factor_configs=configs, log_potentials=rng.normal(size=(2, 6))))
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=[[extra[0], pixels[0, 0]]],
                                        factor_configs=np.array([[0, 0], [1, 1], [1, 2]])))
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[extra[0], extra[2], extra[3], extra[5]]],
      factor_configs=np.array([[0, 0, 0, 0], [1, 2, 1, 2], [0, 1, 1, 0], [1, 1, 0, 2], [0, 2, 0, 1]]),
      log_potentials=rng.normal(size=(1, 5))))
  bp = infer.BP(fg.bp_state, temperature=temperature)
  lead = () if batch is None else (batch,)
  arrays = bp.init(evidence_updates={pixels: rng.gumbel(size=lead + (10, 10, 3)),
                                     hidden: rng.gumbel(size=lead + (8, 8, 17)),
                                     extra: rng.gumbel(size=lead + (6, 4))})
  lp = np.array(arrays.log_potentials, dtype=np.float32)
  lp[rng.integers(0, lp.size, size=4)] = -3e6
  if lp_batched:
    lp = (lp[None, :] + 0.3 * rng.normal(size=(batch, lp.size))).astype(np.float32)
  arrays = BPArrays(log_potentials=lp, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  plan = bp.context.plan
  iters = 7
  got, got_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  # the complete 17 x 3 tables take the nested-loop kernel k_enum_pair_dense: the same operations in the
  # same order as the configuration-major walk (PATH_ENUM_DENSE_PAIR disabled), for every temperature
  plan.disable_paths(plan.PATH_ENUM_DENSE_PAIR)
  cm, cm_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  np.testing.assert_array_equal(got.ftov_msgs, cm.ftov_msgs)
  np.testing.assert_array_equal(got_d, cm_d)
  # ... and, with a 2 ... 4-state side, shared potentials and full sample tiles, k_enum_pair_few (that side
  # in registers): again the same operations in the same order as the nested-loop kernel
  plan.disable_paths(plan.PATH_ENUM_PAIR_FEW)
  dense, dense_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  np.testing.assert_array_equal(got.ftov_msgs, dense.ftov_msgs)
  np.testing.assert_array_equal(got_d, dense_d)
  plan.disable_paths(plan.PATH_ENUM_CONFIG_MAJOR)
  ref, ref_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5,
                                          temperature)
  got_m, ref_m = np.asarray(got.ftov_msgs, np.float64), np.asarray(ref.ftov_msgs, np.float64)
  want_m = np.asarray(want, np.float64).reshape(got_m.shape)
  if temperature == 0.0:
    np.testing.assert_array_equal(got_m, ref_m)
    np.testing.assert_array_equal(got_d, ref_d)
    np.testing.assert_array_equal(got_m, want_m)
    return
  with bp_oracle.precision(np.float64):
    exact, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5,
                                        temperature)
  exact_m = np.asarray(exact, np.float64).reshape(got_m.shape)
  floor = exact_m <= -1e31
  np.testing.assert_array_equal(got_m <= -1e31, floor)
  err_got, err_ref = np.abs(got_m - exact_m)[~floor].max(), np.abs(ref_m - exact_m)[~floor].max()
  err_oracle = np.abs(want_m - exact_m)[~floor].max()
  print(f"vs fp64: configuration-major {err_got:.3g}, list walk {err_ref:.3g}, fp32 oracle {err_oracle:.3g}")
  # on this graph (hidden variables of degree 10 - 18, 7 iterations) the serial fp32 oracle itself is
  # 1e-5 ... 2.4e-5 from the fp64 run, so the device is held to the oracle's own distance from the truth
  # and, directly, to the sum of the two
  assert err_got <= max(1e-5, 2.0 * err_oracle), (err_got, err_oracle)
  np.testing.assert_allclose(got_m[~floor], want_m[~floor], atol=5e-5, rtol=2e-6)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
@pytest.mark.parametrize("batch", [17, 40, 100])
def test_generic_two_pass_binary_difference_storage(temperature, batch):
  """PATH_GENERIC_BIN: an all-binary graph of pairwise + unary EnumFactors that is neither a grid
  nor a dense bipartite block (random edges, a hub variable of degree 30) at full sample tiles keeps
  one float per edge between iterations (k_var_sums_bin / k_enum_pw2_bin / k_enum_unary_bin).
  Expanding is exact: messages, deltas and decodings are bit-identical to the reference-layout
  kernels (path disabled), for zero, shared and batched un-normalised initial messages; and equal
  to the oracle (max-product: bit for bit)."""
  from pgmax_b200.infer.bp_state import BPArrays
  rng = np.random.default_rng(4)
  n = 60
  variables = vgroup.NDVarArray(num_states=2, shape=(n,))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = {(0, j) for j in range(1, 31)}
  while len(pairs) < 150:
    i, j = sorted(rng.integers(0, n, size=2))
    if i != j:
      pairs.add((int(i), int(j)))
  pairs = sorted(pairs)
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=[[variables[i], variables[j]] for i, j in pairs],
      log_potential_matrix=rng.normal(size=(len(pairs), 2, 2))))
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=[[variables[i]] for i in range(0, n, 3)],
                                        factor_configs=np.arange(2)[:, None],
                                        log_potentials=rng.normal(size=(len(range(0, n, 3)), 2))))
  bp = infer.BP(fg.bp_state, temperature=temperature)
  plan = bp.context.plan
  base = bp.init(evidence_updates={variables: rng.gumbel(size=(batch, n, 2))})
  num_msgs = base.ftov_msgs.shape[-1]
  graph = bp_oracle.graph_from_context(bp.context)
  inits = {"zero": base.ftov_msgs,
           "shared": (rng.normal(size=num_msgs) * 2).astype(np.float32),
           "batched": (rng.normal(size=(batch, num_msgs)) * 2).astype(np.float32)}
  for name, msgs in inits.items():
    arrays = BPArrays(log_potentials=base.log_potentials, ftov_msgs=msgs, evidence=base.evidence)
    before = plan.launch_count
    got, got_d = bp.run_with_diffs(arrays, num_iters=6, damping=0.5, temperature=temperature)
    assert plan.launch_count > before
    plan.disable_paths(plan.PATH_GENERIC_BIN)
    ref, ref_d = bp.run_with_diffs(arrays, num_iters=6, damping=0.5, temperature=temperature)
    plan.disable_paths(0)
    np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs, err_msg=name)
    np.testing.assert_array_equal(got_d, ref_d, err_msg=name)
    want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 6, 0.5,
                                            temperature)
    if temperature == 0.0:
      np.testing.assert_array_equal(got.ftov_msgs, np.asarray(want).reshape(got.ftov_msgs.shape), err_msg=name)
    else:
      np.testing.assert_allclose(got.ftov_msgs, np.asarray(want).reshape(got.ftov_msgs.shape), atol=1e-5, err_msg=name)
    for _ in range(2):  # capture + replay
      again, again_d = bp.run_with_diffs(arrays, num_iters=6, damping=0.5, temperature=temperature)
      np.testing.assert_array_equal(again.ftov_msgs, got.ftov_msgs)
      np.testing.assert_array_equal(again_d, got_d)
  # decoding from the variable sums the run leaves behind (pgx_infer_host: PGX_RUN_FINAL_SUMS)
  res = bp.infer_host(arrays, num_iters=6, damping=0.5, marginals=True)
  states, marg, ties = bp.context.decode(got, marginals=True)
  np.testing.assert_array_equal(res["flat_map_states"], states)
  np.testing.assert_array_equal(res["tie_counts"], ties)
  np.testing.assert_array_equal(res["marginals"], marg)
