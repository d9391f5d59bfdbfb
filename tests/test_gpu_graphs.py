"""One launch per run: the second pgx_bp_run call with the same signature captures the launch
sequence of all iterations in a CUDA graph, later calls replay it (include/pgx.h,
PGX_RUN_NO_GRAPH; the reference's loop is one device-side lax.scan, pgmax/infer/bp.py:142-146).
Direct run == capturing run == replay, bit for bit, on every launch path."""

import numpy as np
import pytest
import torch

import models
from oracle import bp_oracle
from pgmax_b200 import infer

pytestmark = pytest.mark.gpu


def _device_arrays(arrays):
  put = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C")).cuda()
  return put(arrays.log_potentials), put(arrays.evidence), put(arrays.ftov_msgs)


def _three_runs(bp, arrays, iters, temperature, with_deltas=True):
  """Same buffers three times: direct, capture + launch, replay.  Returns outputs + counters."""
  plan = bp.context.plan
  lp, ev, msgs = _device_arrays(arrays)
  batch = arrays.batch_size or 1
  out = torch.empty((batch, plan.num_edge_states), dtype=torch.float32, device="cuda")
  deltas = torch.empty((batch, iters), dtype=torch.float32, device="cuda") if with_deltas else None
  stream = torch.cuda.current_stream().cuda_stream
  results, graphs, launches = [], [], []
  for _ in range(3):
    out.fill_(float("nan"))
    g0, l0 = plan.graph_launch_count, plan.launch_count
    plan.bp_run(stream, batch, lp.data_ptr(), lp.ndim == 2, ev.data_ptr(), ev.ndim == 2, msgs.data_ptr(),
                msgs.ndim == 2, out.data_ptr(), deltas.data_ptr() if with_deltas else None, iters, 0.5, temperature)
    torch.cuda.synchronize()
    results.append((out.cpu().numpy().copy(), deltas.cpu().numpy().copy() if with_deltas else None))
    graphs.append(plan.graph_launch_count - g0)
    launches.append(plan.launch_count - l0)
  return results, graphs, launches


def _check(results, graphs, launches):
  assert graphs == [0, 1, 1], graphs            # direct, capture + launch, replay
  assert launches[0] == launches[1] == launches[2] > 0, launches
  for got, got_d in results[1:]:
    np.testing.assert_array_equal(got, results[0][0])
    if got_d is not None:
      np.testing.assert_array_equal(got_d, results[0][1])


@pytest.mark.parametrize("temperature", [0.0, 1.0])
@pytest.mark.parametrize("batch", [None, 3, 40])
def test_graph_replay_ising(batch, temperature):
  """Small pairwise grids: one sample (resident cluster launch inside the graph), small batches
  (pull kernels), and against the oracle."""
  fg, variables, evidence = models.ising_model(n=12, batch=batch)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates={variables: evidence})
  results, graphs, launches = _three_runs(bp, arrays, 20, temperature)
  _check(results, graphs, launches)
  graph = bp_oracle.graph_from_context(bp.context)
  want, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 20, 0.5, temperature)
  np.testing.assert_allclose(results[2][0].reshape(np.asarray(want).shape), want, atol=1e-5)


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_graph_replay_rbm_single_pass_two_streams(temperature):
  """Dense-grid single-pass path incl. the half-batch pipeline (two streams joined by events
  inside the captured graph) and the unary blocks on the auxiliary stream."""
  rng = np.random.default_rng(0)
  W, bh, bv = 0.4 * rng.normal(size=(6, 9)), rng.logistic(size=6), rng.logistic(size=9)
  fg, hidden, visible = models.rbm_model(W, bh, bv)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  for batch in (64, 530):
    arrays = bp.init(evidence_updates={hidden: rng.gumbel(size=(batch, 6, 2)), visible: rng.gumbel(size=(batch, 9, 2))})
    _check(*_three_runs(bp, arrays, 9, temperature))


def test_graph_replay_deconvolution():
  """OR / AND graphs: small batches (generic kernels, one stream) and the fused OR + AND launch
  (full sample tiles + the batch tail on a second stream and a second plan) replay a graph.  With
  the fused launch disabled the OR and AND groups run on two streams of different PRIORITY, which
  a captured graph does not preserve (measured slower, pgx.cu): those runs are enqueued directly.
  Same bits either way."""
  fg, groups = models.deconv_model(im_height=10, im_width=10, n_feat=3, feat_height=3, feat_width=3)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  plan = bp.context.plan
  arrays = bp.init(evidence_updates=models.deconv_evidence(groups, batch=5))
  _check(*_three_runs(bp, arrays, 6, 0.0))
  arrays = bp.init(evidence_updates=models.deconv_evidence(groups, batch=40))
  fused = _three_runs(bp, arrays, 6, 0.0)
  _check(*fused)
  plan.disable_paths(plan.PATH_ORAND_FUSED)
  results, graphs, launches = _three_runs(bp, arrays, 6, 0.0)
  plan.disable_paths(0)
  assert graphs == [0, 0, 0] and launches[0] == launches[2]
  np.testing.assert_array_equal(results[2][0], results[0][0])
  np.testing.assert_array_equal(results[0][0], fused[0][0][0])


def test_graph_replay_rcn_merged_max_product():
  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=6, radii=(2, 4), extra_edges=2, seed=5)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  results, graphs, launches = _three_runs(bp, arrays, 5, 0.0)
  _check(results, graphs, launches)
  graph = bp_oracle.graph_from_context(bp.context)
  want, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 5, 0.5, 0.0)
  np.testing.assert_array_equal(results[2][0][0], want)


def test_graph_goes_stale_when_the_workspace_moves():
  """A cached graph holds workspace pointers: a run with another batch size reallocates the
  workspace, after which the old signature is captured afresh instead of replayed."""
  fg, variables, evidence = models.ising_model(n=10, batch=5)
  bp = infer.BP(fg.bp_state, temperature=1.0)
  plan = bp.context.plan
  arrays = bp.init(evidence_updates={variables: evidence})
  results, graphs, _ = _three_runs(bp, arrays, 8, 1.0, with_deltas=False)
  assert graphs == [0, 1, 1]
  other = bp.init(evidence_updates={variables: np.concatenate([evidence, evidence])})
  bp.run(other, num_iters=3, damping=0.5)          # batch 10: the workspace is reallocated
  lp, ev, msgs = _device_arrays(arrays)
  out = torch.empty((5, plan.num_edge_states), dtype=torch.float32, device="cuda")
  stream = torch.cuda.current_stream().cuda_stream
  for _ in range(3):
    plan.bp_run(stream, 5, lp.data_ptr(), False, ev.data_ptr(), True, msgs.data_ptr(), False, out.data_ptr(), None,
                8, 0.5, 1.0)
  torch.cuda.synchronize()
  np.testing.assert_array_equal(out.cpu().numpy(), results[0][0])


def test_graphs_can_be_switched_off():
  fg, variables, evidence = models.ising_model(n=8, batch=4)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  bp.context.plan.enable_graphs(False)
  arrays = bp.init(evidence_updates={variables: evidence})
  results, graphs, _ = _three_runs(bp, arrays, 6, 0.0)
  assert graphs == [0, 0, 0]
  np.testing.assert_array_equal(results[2][0], results[0][0])
