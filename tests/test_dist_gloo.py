"""Host logic of the multi-GPU modes, world_size 2 and 3 over gloo on CPU (no GPU needed).

The local BP iteration is injected: here an oracle-backed engine on CPU tensors stands in
for the CUDA engine, so what is tested is the partitioning, the halo index arrays and the
exchange protocol of pgmax_b200/dist.py — against the oracle on the UNPARTITIONED graph.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import models
from oracle import bp_oracle
from pgmax_b200 import dist as pdist
from pgmax_b200 import infer


class OracleStepEngine:
  """engine.step / engine.beliefs of StripRunner on top of the NumPy oracle."""

  def __init__(self, flat):
    self.graph = bp_oracle.graph_from_flat(flat)

  def step(self, lp, ev, msgs_in, msgs_out, damping, temperature):
    T = temperature if temperature == 0.0 else np.float32(temperature)
    new, _ = bp_oracle.bp_update(self.graph, msgs_in.numpy(), ev.numpy(),
                                 np.clip(lp.numpy(), -1e6, 1e6), damping, T)
    msgs_out.copy_(torch.from_numpy(new))

  def beliefs(self, ev, msgs):
    return torch.from_numpy(bp_oracle.flat_beliefs(self.graph, msgs.numpy(), ev.numpy()))


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _strip_worker(rank, world, port, n, temperature, iters, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    evidence = np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)
    strip = pdist.ising_strip(n, rank, world)
    runner = pdist.StripRunner(strip, OracleStepEngine(strip.flat), "cpu")
    ev_own = evidence[strip.row0 : strip.row0 + strip.rows].reshape(-1)
    msgs, ev_t = runner.run(ev_own, iters, 0.5, temperature)
    beliefs = runner.beliefs(ev_t, msgs)
    # reference: the whole torus on one "device"
    whole = pdist.ising_strip(n, 0, 1)
    graph = bp_oracle.graph_from_flat(whole.flat)
    want, _ = bp_oracle.run_bp(graph, whole.log_potentials, np.zeros(whole.num_msgs, np.float32),
                               evidence.reshape(-1), iters, 0.5, temperature)
    want_b = bp_oracle.flat_beliefs(graph, want, evidence.reshape(-1))
    lo, hi = strip.global_msg_range
    out[rank] = (float(np.max(np.abs(msgs.numpy() - want[lo:hi]))),
                 float(np.max(np.abs(beliefs.numpy() - want_b[2 * n * strip.row0 : 2 * n * (strip.row0 + strip.rows)]))))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,temperature", [(2, 8, 0.0), (2, 9, 1.0), (3, 7, 1.0)])
def test_row_strips_match_single_graph(world, n, temperature):
  manager = mp.Manager()
  out = manager.dict()
  mp.spawn(_strip_worker, args=(world, _free_port(), n, temperature, 12, out), nprocs=world, join=True)
  assert sorted(out.keys()) == list(range(world))
  for rank in range(world):
    err_m, err_b = out[rank]
    assert err_m <= 2e-6 and err_b <= 4e-6, (rank, err_m, err_b)


def test_single_strip_is_the_reference_ising_graph():
  """world == 1: the array-form generator equals the graph the facade builds."""
  n = 6
  fg, variables, evidence = models.ising_model(n=n)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  g_ref = bp_oracle.graph_from_context(bp.context)
  whole = pdist.ising_strip(n)
  g = bp_oracle.graph_from_flat(whole.flat)
  np.testing.assert_array_equal(g.var_states_for_edge_states, g_ref.var_states_for_edge_states)
  np.testing.assert_array_equal(g.edge_indices_for_edge_states, g_ref.edge_indices_for_edge_states)
  for key in ("factor_configs_indices", "factor_configs_edge_states"):
    np.testing.assert_array_equal(g.inference_arguments["enum"][key],
                                  np.asarray(g_ref.inference_arguments["enum"][key]).reshape(
                                      g.inference_arguments["enum"][key].shape))
  np.testing.assert_allclose(whole.log_potentials, bp.init().log_potentials)


def _batch_worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    fg, variables, evidence = models.ising_model(n=4, batch=5)
    bp = infer.BP(fg.bp_state, temperature=0.0)
    arrays = bp.init(evidence_updates={variables: evidence})
    local = pdist.shard_batch(arrays, world, rank)
    graph = bp_oracle.graph_from_context(bp.context)
    msgs, _ = bp_oracle.run_bp_batched(graph, local.log_potentials, local.ftov_msgs, local.evidence, 5, 0.5, 0.0)
    states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, msgs, local.evidence))
    full = pdist.all_gather_batch(states, 5)
    want_m, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 5, 0.5, 0.0)
    want, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, want_m, arrays.evidence))
    out[rank] = bool(np.array_equal(full.numpy(), want)) and local.log_potentials.ndim == 1
  finally:
    dist.destroy_process_group()


def test_batch_sharding_needs_no_collective_and_gathers_in_order():
  assert [pdist.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
  assert [pdist.shard_bounds(2, 3, r) for r in range(3)] == [(0, 1), (1, 2), (2, 2)]
  with pytest.raises(ValueError):
    pdist.shard_bounds(4, 2, 2)
  manager = mp.Manager()
  out = manager.dict()
  mp.spawn(_batch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  assert out[0] and out[1]


# ---- general factor partition (pdist.partition_flat / PartitionRunner) ----------------------
def _general_graph(which):
  """(FactorGraph, evidence updates) of an irregular graph: EnumFactors ("cut", "ragged"), a random
  network of OR + AND factors on shared leaves ("logical"), a hierarchy of PoolFactors ("pool")."""
  if which == "logical":
    import test_gpu_logical_pull
    fg, groups = test_gpu_logical_pull.random_logical_network(0)
    rng = np.random.default_rng(2)
    return fg, {g: rng.gumbel(size=g.shape + (2,)) * 2.0 for g in groups.values()}
  if which == "pool":
    fg, variables = models.sdlp_pool_model()
    return fg, {variables: np.random.default_rng(4).gumbel(size=(variables.shape[0], 2))}
  if which == "cut":
    fg, bp_state, grid_vars, additional_vars = models.cut_model()
    return fg, None
  rng = np.random.default_rng(5)
  from pgmax_b200 import fgraph, fgroup, vgroup
  num_states = rng.integers(2, 6, size=(14,))
  variables = vgroup.NDVarArray(num_states=num_states, shape=(14,))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = [(int(a), int(b)) for a in range(14) for b in range(a + 1, 14) if rng.random() < 0.3]
  for a, b in pairs:  # ragged numbers of states: one single-factor group per pair
    fg.add_factors(fgroup.PairwiseFactorGroup(
        variables_for_factors=[[variables[a], variables[b]]],
        log_potential_matrix=rng.normal(size=(int(num_states[a]), int(num_states[b])))))
  triple_cfg = np.array([[0, 0, 0], [1, 1, 0], [0, 1, 1], [1, 0, 1]])
  fg.add_factors(fgroup.EnumFactorGroup(
      variables_for_factors=[[variables[0], variables[5], variables[9]], [variables[3], variables[5], variables[12]]],
      factor_configs=triple_cfg, log_potentials=rng.normal(size=(2, 4))))
  return fg, {variables: rng.gumbel(size=(14, int(num_states.max())))}


def _partition_worker(rank, world, port, which, temperature, iters, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    fg, evidence = _general_graph(which)
    bp = infer.BP(fg.bp_state, temperature=temperature)
    arrays = bp.init(evidence_updates=evidence) if evidence is not None else bp.init()
    flat = pdist.flat_from_state(fg.bp_state.fg_state)
    part = pdist.partition_flat(flat, world, rank)
    runner = pdist.PartitionRunner(part, OracleStepEngine(part.flat), "cpu")
    lp = np.asarray(arrays.log_potentials, np.float32)
    ev = np.asarray(arrays.evidence, np.float32)
    msgs = runner.run(lp[part.potential_index], ev[part.var_state_index], iters, 0.5, temperature)
    beliefs = runner.beliefs(ev[part.var_state_index], msgs)
    graph = bp_oracle.graph_from_context(bp.context)
    want, _ = bp_oracle.run_bp(graph, lp, arrays.ftov_msgs, ev, iters, 0.5, temperature)
    want_b = bp_oracle.flat_beliefs(graph, want, ev)
    finite = np.isfinite(want_b[part.var_state_index])
    out[rank] = (float(np.max(np.abs(msgs.numpy() - want[part.msg_index]))),
                 float(np.max(np.abs(beliefs.numpy()[finite] - want_b[part.var_state_index][finite]))),
                 int(part.num_shared), int(part.msg_index.shape[0]))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("world,which,temperature", [(2, "cut", 0.0), (3, "cut", 1.0), (2, "ragged", 1.0),
                                                      (3, "ragged", 0.0), (2, "logical", 0.0), (3, "logical", 0.0),
                                                      (2, "pool", 0.0), (3, "pool", 0.8)])
def test_factor_partition_matches_single_graph(world, which, temperature):
  manager = mp.Manager()
  out = manager.dict()
  mp.spawn(_partition_worker, args=(world, _free_port(), which, temperature, 10, out), nprocs=world, join=True)
  assert sorted(out.keys()) == list(range(world))
  total_msgs = 0
  for rank in range(world):
    err_m, err_b, num_shared, num_msgs = out[rank]
    assert num_shared > 0
    assert err_m <= 1e-5 and err_b <= 2e-5, (rank, err_m, err_b)
    total_msgs += num_msgs
  fg, _ = _general_graph(which)
  assert total_msgs == fg.bp_state.ftov_msgs.value.shape[0]  # every message is owned by exactly one rank


def test_partition_index_arrays_single_rank():
  """world == 1: the part IS the graph (identity index arrays, nothing shared)."""
  fg, _ = _general_graph("ragged")
  flat = pdist.flat_from_state(fg.bp_state.fg_state)
  part = pdist.partition_flat(flat, 1, 0)
  np.testing.assert_array_equal(part.msg_index, np.arange(fg.bp_state.ftov_msgs.value.shape[0]))
  np.testing.assert_array_equal(part.potential_index, np.arange(fg.bp_state.log_potentials.value.shape[0]))
  assert part.num_shared == 0 and part.shared_local_vs.size == 0
  np.testing.assert_array_equal(part.flat.edge_var_start, flat.edge_var_start)
