"""GPU tests of the logical pull path (k_logical_pull_small / _wide + k_var_sums_list): graphs of
OR / AND factors only, batch >= 17 (full sample tiles).  The pull kernels re-derive the sums of
variables with one or two edges in the serial order of k_var_sums, so they must be BIT-identical
to the two-pass kernels (PATH_LOGICAL_PULL disabled) and agree with the oracle."""

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import fgraph, fgroup, infer, vgroup

pytestmark = pytest.mark.gpu


def random_logical_network(seed, and_parents=(1, 4), or_parents=(1, 40)):
  """Leaves (shared by many AND factors: high degree) -> AND -> mid variables (degree 2, some 3)
  -> OR -> outputs (degree 1); plus AND factors on top of OR outputs of an earlier layer."""
  rng = np.random.default_rng(seed)
  n_leaf, n_mid = 12, int(rng.integers(30, 60))
  n_out = int(rng.integers(3, 8))
  leaf = vgroup.NDVarArray(num_states=2, shape=(n_leaf,))
  mid = vgroup.NDVarArray(num_states=2, shape=(n_mid,))
  out = vgroup.NDVarArray(num_states=2, shape=(n_out,))
  top = vgroup.NDVarArray(num_states=2, shape=(2,))
  fg = fgraph.FactorGraph(variable_groups=[leaf, mid, out, top])
  ands = []
  for i in range(n_mid):
    k = int(rng.integers(and_parents[0], and_parents[1] + 1))
    ands.append([leaf[int(j)] for j in rng.choice(n_leaf, size=k, replace=False)] + [mid[i]])
  ors = []
  pool = list(rng.permutation(n_mid))
  for o in range(n_out):
    k = int(min(len(pool), rng.integers(or_parents[0], or_parents[1] + 1))) if o < n_out - 1 else len(pool)
    k = max(k, 1) if pool else 0
    take, pool = pool[:k], pool[k:]
    extra = [int(j) for j in rng.choice(n_mid, size=2, replace=False) if int(j) not in take][:1]  # degree-3 mids
    if take or extra:
      ors.append([mid[int(j)] for j in list(take) + extra] + [out[o]])
  fg.add_factors(fgroup.ORFactorGroup(ors))
  # uniform AND groups per parent count, plus the outputs feeding two more ANDs
  by_k = {}
  for v in ands:
    by_k.setdefault(len(v), []).append(v)
  for k in sorted(by_k):
    fg.add_factors(fgroup.ANDFactorGroup(by_k[k]))
  fg.add_factors(fgroup.ANDFactorGroup([[out[0], out[1 % n_out], top[0]], [out[0], leaf[0], top[1]]]))
  groups = dict(leaf=leaf, mid=mid, out=out, top=top)
  return fg, groups


def _evidence(groups, batch, seed):
  rng = np.random.default_rng(seed)
  return {g: rng.gumbel(size=(batch,) + g.shape + (2,)).astype(np.float32) * 2.0 for g in groups.values()}


def _close_to_oracle(got, want, temperature, atol=2e-5):
  """Max-product: tight.  Sum-product with many parents: the reference's closed form takes
  logminusexp(L, Sb) of two sums of magnitude ~1e2 that differ by ~1e-3, behind an eps cut-off
  (logical.py:673-737): a 1-ulp difference between two correct expf / log1pf implementations moves
  that difference by 1 % or flips the cut-off, so a few per cent of the messages differ at the
  1e-3 level between ANY two fp32 evaluations (the GPU paths agree with each other bit for bit,
  asserted by the callers).  Bound the fraction of such entries and their typical size."""
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  assert not np.isnan(got).any()
  if temperature == 0.0:
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=atol)
    return
  floor_g, floor_w = got <= -1e31, ~np.isfinite(want) | (want <= -1e31)
  both = ~floor_g & ~floor_w
  err = np.abs(got[both] - want[both])
  bad = (err > atol + 2e-6 * np.abs(want[both])).sum() + (floor_g != floor_w).sum()
  assert bad / got.size < 0.10, bad / got.size
  assert np.median(err) < atol, np.median(err)


def _run(bp, arrays, iters, temperature, mask):
  plan = bp.context.plan
  plan.disable_paths(mask)
  out = bp.run_with_diffs(arrays, num_iters=iters, damping=0.5, temperature=temperature)
  plan.disable_paths(0)
  return out


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("temperature", [0.0, 0.3, 0.8])
@pytest.mark.parametrize("batch", [17, 40])
def test_pull_equals_two_pass_and_oracle(seed, temperature, batch):
  fg, groups = random_logical_network(seed)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=_evidence(groups, batch, seed))
  plan = bp.context.plan
  launches = plan.launch_count
  got, got_d = _run(bp, arrays, 6, temperature, 0)
  kinds = plan.launch_count - launches
  ref, ref_d = _run(bp, arrays, 6, temperature, plan.PATH_LOGICAL_PULL)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  assert kinds > 0
  # the pull kernels on the reference layout (two floats per edge) instead of binary-difference storage
  full, full_d = _run(bp, arrays, 6, temperature, plan.PATH_LOGICAL_BIN)
  np.testing.assert_array_equal(got.ftov_msgs, full.ftov_msgs)
  np.testing.assert_array_equal(got_d, full_d)
  # the small-factor kernel with its wiring read from global memory per warp iteration
  # instead of staged per CTA
  unstaged, unstaged_d = _run(bp, arrays, 6, temperature, plan.PATH_STAGED_WIRING)
  np.testing.assert_array_equal(got.ftov_msgs, unstaged.ftov_msgs)
  np.testing.assert_array_equal(got_d, unstaged_d)
  # single-launch wide kernel, everything on one stream
  one, one_d = _run(bp, arrays, 6, temperature, plan.PATH_WIDE_SPLIT | plan.PATH_AUX_STREAM)
  np.testing.assert_array_equal(got.ftov_msgs, one.ftov_msgs)
  np.testing.assert_array_equal(got_d, one_d)
  graph = bp_oracle.graph_from_context(bp.context)
  if temperature == 0.0:
    want, want_d = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                                            6, 0.5, temperature)
    _close_to_oracle(got.ftov_msgs, want, temperature)
    np.testing.assert_allclose(got_d, want_d, rtol=1e-5, atol=2e-5)
  else:
    # ONE iteration from the state the GPU reached (no propagation of flipped cut-offs)
    one, _ = _run(bp, got, 1, temperature, 0)
    want, _ = bp_oracle.run_bp_batched(graph, got.log_potentials, got.ftov_msgs, got.evidence, 1, 0.5, temperature)
    _close_to_oracle(one.ftov_msgs, want, temperature)


@pytest.mark.parametrize("temperature", [0.0, 0.5])
def test_pull_deconvolution_batch_40(temperature):
  """Small deconvolution graph (uniform 2-parent ANDs, ORs with up to 18 parents; S and W are
  the high-degree variables, SW and X are pulled), 40 images, with shared and per-sample evidence."""
  fg, groups = models.deconv_model(im_height=9, im_width=8, n_feat=2, feat_height=3, feat_width=3)
  evidence = models.deconv_evidence(groups, batch=40)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=evidence)
  plan = bp.context.plan
  got, got_d = _run(bp, arrays, 12, temperature, 0)
  ref, ref_d = _run(bp, arrays, 12, temperature, plan.PATH_LOGICAL_PULL)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  full, full_d = _run(bp, arrays, 12, temperature, plan.PATH_LOGICAL_BIN)
  np.testing.assert_array_equal(got.ftov_msgs, full.ftov_msgs)
  np.testing.assert_array_equal(got_d, full_d)
  graph = bp_oracle.graph_from_context(bp.context)
  if temperature == 0.0:
    want, _ = bp_oracle.run_bp_batched(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                                       12, 0.5, temperature)
    np.testing.assert_allclose(got.ftov_msgs, want, rtol=2e-6, atol=2e-4)  # messages reach 230
  else:
    one, _ = _run(bp, got, 1, temperature, 0)
    want, _ = bp_oracle.run_bp_batched(graph, got.log_potentials, got.ftov_msgs, got.evidence, 1, 0.5, temperature)
    _close_to_oracle(one.ftov_msgs, want, temperature, atol=2e-4)


@pytest.mark.parametrize("temperature", [0.0, 0.5])
@pytest.mark.parametrize("batch", [33, 36])
def test_batch_tail_runs_beside_the_full_tiles(temperature, batch):
  """PATH_TAIL_SPLIT: the <= 8 samples beyond the last full tile of 32 go through a second plan
  (generic kernels) on its own stream.  Bit-identical to the unsplit run - messages, deltas,
  batched initial messages (second run), launch count includes the tail's kernels."""
  fg, groups = models.deconv_model(im_height=9, im_width=8, n_feat=2, feat_height=3, feat_width=3)
  evidence = models.deconv_evidence(groups, batch=batch)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=evidence)
  plan = bp.context.plan
  # (the fused OR + AND launch takes a partial tile instead of splitting the tail: the split is
  # the separate kernels' mechanism)
  sep = plan.PATH_ORAND_FUSED
  before = plan.launch_count
  ref, ref_d = _run(bp, arrays, 7, temperature, sep | plan.PATH_TAIL_SPLIT)
  unsplit_launches = plan.launch_count - before
  before = plan.launch_count
  got, got_d = _run(bp, arrays, 7, temperature, sep)
  split_launches = plan.launch_count - before
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  ref2, ref2_d = _run(bp, ref, 5, temperature, sep | plan.PATH_TAIL_SPLIT)
  got2, got2_d = _run(bp, got, 5, temperature, sep)
  np.testing.assert_array_equal(got2.ftov_msgs, ref2.ftov_msgs)
  np.testing.assert_array_equal(got2_d, ref2_d)
  assert split_launches > unsplit_launches  # the tail's kernels are counted
  # decode of the split run's messages against the oracle's decode of the same messages
  graph = bp_oracle.graph_from_context(bp.context)
  states, _, _ = bp.context.decode(got2)
  want_states, _, _ = bp_oracle.decode_flat(graph, bp_oracle.flat_beliefs(graph, got2.ftov_msgs, got2.evidence))
  np.testing.assert_array_equal(states, want_states)


@pytest.mark.parametrize("temperature", [0.0, 0.5])
@pytest.mark.parametrize("batch", [32, 40, 70])
def test_fused_or_and_launch_is_bit_identical(temperature, batch):
  """k_or_and_fused (one launch per iteration for graphs whose OR parents are the degree-2
  children of two-parent AND factors: the deconvolution model) against the separate AND / OR
  reduce / OR emit kernels (PATH_ORAND_FUSED disabled): same helpers, same order of operations,
  so every message and delta is identical; partial last tiles (batch 70 = 2 tiles + a tail of 6,
  batch 40 = one tile + a tail of 8) included."""
  fg, groups = models.deconv_model(im_height=9, im_width=8, n_feat=2, feat_height=3, feat_width=3)
  evidence = models.deconv_evidence(groups, batch=batch)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=evidence)
  plan = bp.context.plan
  before = plan.launch_count
  got, got_d = _run(bp, arrays, 9, temperature, 0)
  fused_launches = plan.launch_count - before
  before = plan.launch_count
  ref, ref_d = _run(bp, arrays, 9, temperature, plan.PATH_ORAND_FUSED)
  separate_launches = plan.launch_count - before
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  assert fused_launches < separate_launches
  # un-normalised batched initial messages, no tail split
  rng = np.random.default_rng(5)
  from pgmax_b200.infer.bp_state import BPArrays
  init = BPArrays(log_potentials=arrays.log_potentials, evidence=arrays.evidence,
                  ftov_msgs=rng.normal(size=(batch, arrays.ftov_msgs.shape[-1])).astype(np.float32))
  a, a_d = _run(bp, init, 4, temperature, plan.PATH_TAIL_SPLIT)
  b, b_d = _run(bp, init, 4, temperature, plan.PATH_TAIL_SPLIT | plan.PATH_ORAND_FUSED)
  np.testing.assert_array_equal(a.ftov_msgs, b.ftov_msgs)
  np.testing.assert_array_equal(a_d, b_d)


@pytest.mark.parametrize("temperature", [0.0, 0.5])
@pytest.mark.parametrize("batch,pack", [(20, 1), (36, 8), (39, 4), (44, 2), (80, 2), (100, 8)])
def test_fused_or_and_packed_batch_tail(temperature, batch, pack):
  """A last tile of <= 16 samples runs as k_or_and_fused<kPack> - 8 / 4 / 2 OR factors per CTA,
  lane = (factor slot, sample) - beside the full tiles (pgx.cu, launch_f2v).  63 OR factors (not
  a multiple of the pack: the last CTA has idle factor slots).  Bit-identical to the whole-tile
  launch (PATH_TAIL_SPLIT disabled: every tile through kPack = 1) and to the separate kernels;
  the CUDA graph replays both launches."""
  del pack  # documents which variant the batch selects
  fg, groups = models.deconv_model(im_height=7, im_width=9, n_feat=2, feat_height=3, feat_width=3)
  bp = infer.BP(fg.bp_state, temperature=temperature)
  arrays = bp.init(evidence_updates=models.deconv_evidence(groups, batch=batch))
  plan = bp.context.plan
  got, got_d = _run(bp, arrays, 7, temperature, 0)
  whole, whole_d = _run(bp, arrays, 7, temperature, plan.PATH_TAIL_SPLIT)
  np.testing.assert_array_equal(got.ftov_msgs, whole.ftov_msgs)
  np.testing.assert_array_equal(got_d, whole_d)
  sep, sep_d = _run(bp, arrays, 7, temperature, plan.PATH_ORAND_FUSED | plan.PATH_TAIL_SPLIT)
  np.testing.assert_array_equal(got.ftov_msgs, sep.ftov_msgs)
  np.testing.assert_array_equal(got_d, sep_d)
  again, again_d = _run(bp, arrays, 7, temperature, 0)   # second / third call: captured graph
  again, again_d = _run(bp, arrays, 7, temperature, 0)
  np.testing.assert_array_equal(got.ftov_msgs, again.ftov_msgs)
  np.testing.assert_array_equal(got_d, again_d)


def test_fused_or_and_full_size_deconvolution():
  """configs[2] graph (28 x 28, 95 220 AND + 784 OR factors of up to 180 parents), 32 images,
  3 max-product iterations: fused launch == separate kernels bit for bit, and the oracle for two
  of the samples."""
  fg, groups = models.deconv_model()
  evidence = models.deconv_evidence(groups, batch=32)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates=evidence)
  plan = bp.context.plan
  got, got_d = _run(bp, arrays, 3, 0.0, 0)
  ref, ref_d = _run(bp, arrays, 3, 0.0, plan.PATH_ORAND_FUSED)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  # the W variables (529 edges each) through one warp per variable instead of the cooperative
  # shared-memory gather (k_var_sums_big_bin): the same additions in the same order
  ref, ref_d = _run(bp, arrays, 3, 0.0, plan.PATH_VARSUM_COOP)
  np.testing.assert_array_equal(got.ftov_msgs, ref.ftov_msgs)
  np.testing.assert_array_equal(got_d, ref_d)
  graph = bp_oracle.graph_from_context(bp.context)
  for b in (0, 31):
    want, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence[b], 3, 0.5, 0.0)
    np.testing.assert_allclose(np.asarray(got.ftov_msgs)[b], want, rtol=2e-6, atol=2e-4)
