"""Parity at the sizes BASELINE.json names (or the largest the oracle finishes in seconds),
one test per config; the RBM config (configs[1]) is in test_gpu_parity.py
(test_rbm_full_size_fused_properties)."""

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import dist as pdist
from pgmax_b200 import infer

pytestmark = pytest.mark.gpu


def test_config0_ising_50x50_1000_iterations():
  """configs[0] exactly: 50x50 torus, 1000 iterations, damping 0.5, T = 0.05 — runs as one
  resident cluster launch; messages within 1e-5 of the oracle after all 1000 iterations,
  MAP identical up to counted ties, deltas decay identically."""
  fg, variables, evidence = models.ising_model(n=50)
  bp = infer.BP(fg.bp_state, temperature=0.05)
  arrays = bp.init(evidence_updates={variables: evidence})
  got, got_d = bp.run_with_diffs(arrays, num_iters=1000, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                                  1000, 0.5, 0.05)
  np.testing.assert_allclose(got.ftov_msgs, want, atol=1e-5)
  # per-iteration max|delta|: before convergence the trajectory amplifies the 1e-7-level
  # difference of the ex2/lg2 softplus, so the mid-run deltas agree to 1e-4,
  # the converged tail (and the final messages above) to 1e-5
  np.testing.assert_allclose(got_d, want_d, atol=1e-4)
  np.testing.assert_allclose(got_d[-200:], want_d[-200:], atol=1e-5)
  states, marg, ties = bp.context.decode(got, marginals=True)
  w_states, w_marg, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, want, arrays.evidence))
  assert int(ties) == 0
  np.testing.assert_array_equal(states, w_states)
  np.testing.assert_allclose(marg, w_marg, atol=1e-5)


def test_config2_deconvolution_28x28_full_graph():
  """configs[2] graph at full size (95 220 AND + 784 OR factors, up to 180 parents), 3 images,
  6 max-product iterations against the oracle.  Messages reach |logit(1e-100)| = 230 (one fp32
  ulp = 1.5e-5), hence the relative tolerance."""
  fg, groups = models.deconv_model()
  evidence = models.deconv_evidence(groups, batch=3)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  got, got_d = bp.run_with_diffs(arrays, num_iters=6, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs,
                                          arrays.evidence, 6, 0.5, 0.0)
  np.testing.assert_allclose(got.ftov_msgs, want, rtol=2e-6, atol=2e-4)
  np.testing.assert_allclose(got_d, want_d, rtol=1e-5, atol=2e-4)
  states, _, _ = bp.context.decode(got)
  beliefs = bp_oracle.flat_beliefs(graph, want, arrays.evidence)
  w_states, _, _ = bp_oracle.decode_flat(graph, beliefs)
  near_tie = np.abs(beliefs.reshape(3, -1, 2)[..., 0] - beliefs.reshape(3, -1, 2)[..., 1]) < 1e-3
  assert np.array_equal(states[~near_tie], w_states[~near_tie])


def test_config3_rcn_625_state_factors():
  """configs[3] factor shape at full size: 625-state variables, perturb radii 2 and 5 (14 161
  and 60 025 valid configurations per factor), one model of 8 variables, 30 max-product
  iterations: bit-exact messages (max is order-independent) and identical MAP."""
  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=8, radii=(2, 5), extra_edges=3, seed=3)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  got = bp.run(arrays, num_iters=30, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, 30, 0.5, 0.0)
  np.testing.assert_array_equal(got.ftov_msgs, want)
  states, _, ties = bp.context.decode(got)
  w_states, _, w_ties = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, want, arrays.evidence))
  np.testing.assert_array_equal(states, w_states)
  assert int(ties) == int(w_ties)
  # the per-group kernel (shared-memory atomics) instead of the merged balanced launch
  plan = bp.context.plan
  plan.disable_paths(plan.PATH_MERGED_MAX)
  np.testing.assert_array_equal(bp.run(arrays, num_iters=30, damping=0.5).ftov_msgs, want)
  plan.disable_paths(0)


def test_config3_rcn_random_potentials_round_ordered_copy():
  """Non-zero potentials (a few beyond the +-1e6 clip) through the merged launch: the default
  path reads a round-ordered, clipped copy of the potentials made once per run; it must agree
  bit for bit with the gather path (copy disabled, or fewer than 3 iterations) and the oracle."""
  from pgmax_b200.infer.bp_state import BPArrays
  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=6, radii=(2, 4, 7), extra_edges=2, seed=11)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  rng = np.random.default_rng(7)
  lp = (rng.normal(size=arrays.log_potentials.shape) * 2.0).astype(np.float32)
  lp[rng.integers(0, lp.size, size=50)] = 3e6
  lp[rng.integers(0, lp.size, size=50)] = -np.inf
  arrays = BPArrays(log_potentials=lp, ftov_msgs=arrays.ftov_msgs, evidence=arrays.evidence)
  graph = bp_oracle.graph_from_context(bp.context)
  plan = bp.context.plan
  for iters in (2, 6):
    want, want_d = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5, 0.0)
    got, got_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5)
    np.testing.assert_array_equal(got.ftov_msgs, want)
    np.testing.assert_array_equal(got_d, want_d)
    plan.disable_paths(plan.PATH_PERM_POTENTIALS)
    ref = bp.run(arrays, num_iters=iters, damping=0.5)
    plan.disable_paths(0)
    np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)


def test_config3_rcn_unnormalised_initial_messages():
  """One sample, 625-state edges, random (un-normalised, some below the -1e32 clip) initial
  messages: the input normalisation (bp.py:92-96) runs as a warp per edge
  (k_normalize_edges_warp); bit-identical to the oracle after 1 and 4 iterations."""
  from pgmax_b200.infer.bp_state import BPArrays
  fg, groups, evidence = models.rcn_model(num_models=1, num_vars=6, radii=(2, 4), extra_edges=2, seed=13)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  rng = np.random.default_rng(3)
  msgs = (rng.normal(size=arrays.ftov_msgs.shape) * 3.0).astype(np.float32)
  msgs[rng.integers(0, msgs.size, size=40)] = -3e33
  arrays = BPArrays(log_potentials=arrays.log_potentials, ftov_msgs=msgs, evidence=arrays.evidence)
  graph = bp_oracle.graph_from_context(bp.context)
  for iters in (1, 4):
    want, want_d = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence, iters, 0.5, 0.0)
    got, got_d = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5)
    np.testing.assert_array_equal(got.ftov_msgs, want)
    np.testing.assert_array_equal(got_d, want_d)


def test_config3_rcn_batched_merged_launch():
  """Three samples (different evidence) of a two-model RCN-shaped graph through the merged
  max-product launch: every sample bit-identical to its own single-sample oracle run."""
  fg, groups, evidence = models.rcn_model(num_models=2, num_vars=6, radii=(2, 3, 8), extra_edges=2, seed=5)
  rng = np.random.default_rng(0)
  batched = {vg: np.stack([np.where(rng.random(ev.shape) < 0.05, 1.0, -1.0) for _ in range(3)])
             for vg, ev in evidence.items()}
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=batched)
  got, got_d = bp.run_with_diffs(arrays, num_iters=5, damping=0.5)
  graph = bp_oracle.graph_from_context(bp.context)
  want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                                          5, 0.5, 0.0)
  np.testing.assert_array_equal(got.ftov_msgs, want)
  np.testing.assert_array_equal(got_d, want_d)


def test_config4_ising_1024_sum_product_short_horizon():
  """configs[4] shape at 1024 x 1024 (the oracle needs ~1 s per iteration here; 8192^2 is
  covered by the strip-vs-single-graph tests): 3 sum-product iterations, T = 1."""
  n, iters = 1024, 3
  evidence = np.random.default_rng(0).gumbel(size=(n, n, 2)).astype(np.float32)
  strip = pdist.ising_strip(n)
  runner = pdist.StripRunner(strip, pdist.PgxStepEngine(strip.flat, "cuda:0"), "cuda:0")
  msgs, _ = runner.run(evidence.reshape(-1), iters, 0.5, 1.0)
  graph = bp_oracle.graph_from_flat(strip.flat)
  want, _ = bp_oracle.run_bp(graph, strip.log_potentials, np.zeros(strip.num_msgs, np.float32),
                             evidence.reshape(-1), iters, 0.5, 1.0)
  np.testing.assert_allclose(msgs.cpu().numpy(), want, atol=1e-5)
