"""GPU tests of the index-free lattice path (k_lattice): every path of pgx_bp_run computes the
same arithmetic in the same order, so the lattice kernel must be BIT-identical to the two-pass
path (k_var_sums + k_enum_pw2) for max-product and sum-product, and bit-identical to the
oracle for max-product, on tori and on open strips with a ghost row, at sizes that are not
multiples of the tile."""

import numpy as np
import pytest
import torch

from oracle import bp_oracle
from pgmax_b200 import _native
from pgmax_b200 import dist as pdist

pytestmark = pytest.mark.gpu


def lattice_flat(rows, cols, torus=True):
  """rows x cols lattice laid out as examples/ising_model.ipynb cell 12 (variable (l, j) owns the
  vertical factor to (l + 1, j) and the horizontal factor to (l, j + 1)); torus=False: one more,
  factor-less, ghost row below (the local graph of a row strip, pgmax_b200/dist.py)."""
  ll, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
  v0 = ll * cols + jj
  below = ((ll + 1) % rows if torus else ll + 1) * cols + jj
  right = ll * cols + (jj + 1) % cols
  edge_var = np.stack([v0, below, v0, right], axis=-1).reshape(-1)
  num_factors = 2 * rows * cols
  num_vars = (rows if torus else rows + 1) * cols
  return _native.FlatGraph(
      var_num_states=np.full((num_vars,), 2, dtype=np.int32),
      edge_var_start=(2 * edge_var).astype(np.int32),
      edge_num_states=np.full((2 * num_factors,), 2, dtype=np.int32),
      num_potentials=4 * num_factors,
      enum_blocks=[_native.FlatEnumBlock(
          num_factors=num_factors, factor_configs=np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.int32),
          first_edge=0, first_potential=0)])


def run_plan(plan, lp, ev, msgs0, iters, temperature, with_deltas=True):
  dev = torch.device("cuda", 0)
  t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
  d_lp, d_ev, d_in = t(lp), t(ev), t(msgs0)
  out = torch.empty_like(d_in)
  deltas = torch.zeros(iters, dtype=torch.float32, device=dev) if with_deltas else None
  plan.bp_run(torch.cuda.current_stream(dev).cuda_stream, 1, d_lp.data_ptr(), False, d_ev.data_ptr(), False,
              d_in.data_ptr(), False, out.data_ptr(), deltas.data_ptr() if with_deltas else None, iters, 0.5,
              temperature)
  torch.cuda.synchronize()
  return out.cpu().numpy(), (deltas.cpu().numpy() if with_deltas else None)


@pytest.mark.parametrize("rows,cols,torus", [(2, 2, True), (5, 7, True), (16, 64, True), (17, 65, True),
                                              (40, 130, True), (50, 50, True), (1, 9, False), (7, 33, False),
                                              (33, 70, False)])
@pytest.mark.parametrize("temperature", [0.0, 0.7])
def test_lattice_bit_identical_to_two_pass_and_oracle(rows, cols, torus, temperature):
  flat = lattice_flat(rows, cols, torus)
  plan = _native.Plan(flat)
  assert plan.is_lattice
  rng = np.random.default_rng(rows * 1000 + cols)
  lp = rng.normal(size=flat.num_potentials).astype(np.float32)
  ev = rng.gumbel(size=2 * flat.var_num_states.shape[0]).astype(np.float32)
  msgs0 = rng.normal(size=8 * rows * cols).astype(np.float32)
  iters = 7
  plan.disable_paths(plan.PATH_RESIDENT)  # small graphs would otherwise take the resident kernel
  launches = plan.launch_count
  got, got_d = run_plan(plan, lp, ev, msgs0, iters, temperature)
  assert plan.launch_count - launches == iters + 1  # normalise + one k_lattice per iteration
  plan.disable_paths(plan.PATH_LATTICE | plan.PATH_RESIDENT | plan.PATH_PULL)
  ref, ref_d = run_plan(plan, lp, ev, msgs0, iters, temperature)
  np.testing.assert_array_equal(got, ref)
  np.testing.assert_array_equal(got_d, ref_d)
  graph = bp_oracle.graph_from_flat(flat)
  want, want_d = bp_oracle.run_bp(graph, lp, msgs0, ev, iters, 0.5, temperature)
  if temperature == 0.0:
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got_d, want_d)
  else:
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_lattice_not_detected_on_other_graphs():
  """A lattice with one vertical edge rewired, and the strip graph of a 2-rank split (detected as
  an open lattice), to pin the detector."""
  flat = lattice_flat(6, 8, True)
  assert _native.Plan(flat).is_lattice
  broken = lattice_flat(6, 8, True)
  broken.edge_var_start[1] = 2 * 17  # factor 0 now joins variable 0 with variable 17
  assert not _native.Plan(broken).is_lattice
  strip = pdist.ising_strip(12, rank=1, world=3)
  assert _native.Plan(strip.flat).is_lattice


def test_lattice_large_torus_properties():
  """2048 x 2048 torus (33.5 M edge-states; the oracle would need ~4 s per iteration): the lattice
  path against the two-pass path, bit for bit, 3 sum-product iterations."""
  n = 2048
  flat = lattice_flat(n, n, True)
  plan = _native.Plan(flat)
  rng = np.random.default_rng(1)
  lp = np.tile(np.float32(0.8) * np.array([1, -1, -1, 1], np.float32), 2 * n * n)
  ev = rng.gumbel(size=2 * n * n).astype(np.float32)
  msgs0 = np.zeros(8 * n * n, np.float32)
  got, got_d = run_plan(plan, lp, ev, msgs0, 3, 1.0)
  plan.disable_paths(plan.PATH_LATTICE | plan.PATH_RESIDENT | plan.PATH_PULL)
  ref, ref_d = run_plan(plan, lp, ev, msgs0, 3, 1.0)
  np.testing.assert_array_equal(got, ref)
  np.testing.assert_array_equal(got_d, ref_d)
